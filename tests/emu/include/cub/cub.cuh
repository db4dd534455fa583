// TEST INFRASTRUCTURE (tests/emu): the two cub entry points the library uses, as std algorithms with cub's calling convention
// (a first call with a null scratch pointer reports the scratch size).
#pragma once
#include <cuda_runtime.h>
#include <numeric>
namespace cub {
struct DeviceRadixSort {
    template<typename K, typename V>
    static cudaError_t SortPairs(void *tmp, size_t &bytes, const K *kin, K *kout, const V *vin, V *vout, int n, int begin_bit = 0, int end_bit = int(sizeof(K) * 8), cudaStream_t = nullptr) {
        if (!tmp) { bytes = 256; return cudaSuccess; }
        std::vector<int> idx(size_t(n > 0 ? n : 0));
        std::iota(idx.begin(), idx.end(), 0);
        const int width = end_bit - begin_bit;
        const unsigned long long mask = width >= 64 ? ~0ull : ((1ull << width) - 1ull);
        auto key = [&](int i) { return (static_cast<unsigned long long>(kin[i]) >> begin_bit) & mask; };
        std::stable_sort(idx.begin(), idx.end(), [&](int a, int b) { return key(a) < key(b); });
        std::vector<K> ks(idx.size()); std::vector<V> vs(idx.size());
        for (size_t i = 0; i < idx.size(); ++i) { ks[i] = kin[idx[i]]; vs[i] = vin[idx[i]]; }
        for (size_t i = 0; i < idx.size(); ++i) { kout[i] = ks[i]; vout[i] = vs[i]; }
        return cudaSuccess;
    }
};
struct DeviceScan {
    template<typename I, typename O>
    static cudaError_t ExclusiveSum(void *tmp, size_t &bytes, const I *in, O *out, int n, cudaStream_t = nullptr) {
        if (!tmp) { bytes = 256; return cudaSuccess; }
        O acc = 0;
        for (int i = 0; i < n; ++i) { const O v = O(in[i]); out[i] = acc; acc = O(acc + v); }
        return cudaSuccess;
    }
};
}
