// TEST INFRASTRUCTURE (tests/emu): the fiber scheduler behind tests/emu/include/cuda_runtime.h.
#include <cuda_runtime.h>

namespace emu {
Block *g_block = nullptr;
Fiber *g_cur = nullptr;
static void (*g_entry)(void *) = nullptr;
static void *g_closure = nullptr;
static constexpr size_t STACK = 256 * 1024;

static void trampoline() {
    g_entry(g_closure);
    g_cur->done = true;
    swapcontext(&g_cur->ctx, &g_block->sched);
}

// all live lanes named in the mask have arrived at the same collective?  then compute every lane's result
static bool try_resolve(Block &b, unsigned warp, unsigned mask, op_kind op) {
    const unsigned base = warp * 32, n = unsigned(b.fibers.size());
    for (unsigned l = 0; l < 32; ++l) {
        if (!(mask >> l & 1u) || base + l >= n) continue;
        Fiber &f = b.fibers[base + l];
        if (f.done) continue;                                   // exited lanes do not take part
        if (!f.pending || f.resolved || f.op != op || f.mask != mask) return false;
    }
    auto live = [&](unsigned l) { return (mask >> l & 1u) && base + l < n && !b.fibers[base + l].done; };
    unsigned long long ballot = 0, any = 0, all = 1, rmax = 0, rmin = ~0ull;
    for (unsigned l = 0; l < 32; ++l) if (live(l)) {
        const auto v = b.fibers[base + l].val;
        if (v) ballot |= 1ull << l;
        any |= v ? 1 : 0; all &= v ? 1 : 0;
        rmax = std::max(rmax, v & 0xFFFFFFFFull); rmin = std::min(rmin, v & 0xFFFFFFFFull);
    }
    for (unsigned l = 0; l < 32; ++l) if (live(l)) {
        Fiber &f = b.fibers[base + l];
        switch (op) {
        case OP_SHFL: {
            const int mode = f.arg >> 8, a = f.arg & 0xFF;
            int src = int(l);
            if (mode == 0) src = a; else if (mode == 1) src = int(l) - a; else if (mode == 2) src = int(l) + a; else src = int(l) ^ a;
            if (src < 0 || src > 31 || !live(unsigned(src))) src = int(l);      // out of range / inactive source: own value
            f.result = b.fibers[base + unsigned(src)].val;
            break;
        }
        case OP_BALLOT: f.result = ballot; break;
        case OP_ANY: f.result = any; break;
        case OP_ALL: f.result = all; break;
        case OP_MATCH32: case OP_MATCH64: {
            unsigned long long m = 0;
            for (unsigned k = 0; k < 32; ++k) if (live(k) && b.fibers[base + k].val == f.val) m |= 1ull << k;
            f.result = m;
            break;
        }
        case OP_RMAX: f.result = rmax; break;
        case OP_RMIN: f.result = rmin; break;
        case OP_SYNC: f.result = 0; break;
        }
    }
    for (unsigned l = 0; l < 32; ++l) if (live(l)) { Fiber &f = b.fibers[base + l]; f.resolved = true; f.wait = W_NONE; }
    return true;
}

unsigned long long collective(op_kind op, unsigned mask, unsigned long long val, int arg) {
    Fiber &f = *g_cur;
    const unsigned lane = f.tid.x & 31u, warp = f.tid.x >> 5;
    if (!(mask >> lane & 1u)) { std::fprintf(stderr, "emu: lane %u calls a collective with mask %08x that does not name it\n", lane, mask); std::abort(); }
    f.pending = true; f.resolved = false; f.op = op; f.mask = mask; f.val = val; f.arg = arg;
    if (!try_resolve(*g_block, warp, mask, op)) { f.wait = W_WARP; yield(); }
    f.pending = false;
    return f.result;
}

void run_grid(dim3 grid, dim3 block, size_t smem, void (*entry)(void *), void *closure) {
    if (g_block) { std::fputs("emu: nested launch\n", stderr); std::abort(); }
    const unsigned nthreads = block.x * block.y * block.z;
    static std::vector<char *> stacks;
    while (stacks.size() < nthreads) stacks.push_back(static_cast<char *>(std::aligned_alloc(64, STACK)));
    std::vector<char> dyn(smem + 64);
    Block b;
    b.bdim = block; b.gdim = grid;
    b.dyn_smem = reinterpret_cast<void *>((reinterpret_cast<uintptr_t>(dyn.data()) + 63) & ~uintptr_t(63));
    g_entry = entry; g_closure = closure;
    for (unsigned bz = 0; bz < grid.z; ++bz) for (unsigned by = 0; by < grid.y; ++by) for (unsigned bx = 0; bx < grid.x; ++bx) {
        b.bid = {bx, by, bz};
        b.fibers.assign(nthreads, Fiber{});
        g_block = &b;
        for (unsigned t = 0; t < nthreads; ++t) {
            Fiber &f = b.fibers[t];
            f.tid = {t % block.x, (t / block.x) % block.y, t / (block.x * block.y)};
            getcontext(&f.ctx);
            f.ctx.uc_stack.ss_sp = stacks[t]; f.ctx.uc_stack.ss_size = STACK; f.ctx.uc_link = &b.sched;
            makecontext(&f.ctx, trampoline, 0);
        }
        for (;;) {
            bool ran = false, all_done = true;
            for (unsigned t = 0; t < nthreads; ++t) {
                Fiber &f = b.fibers[t];
                if (f.done) continue;
                all_done = false;
                if (f.wait != W_NONE) continue;
                g_cur = &f;
                swapcontext(&b.sched, &f.ctx);
                ran = true;
            }
            if (all_done) break;
            // barrier: every live fiber has arrived
            bool at_barrier = true, any_live = false;
            for (const Fiber &f : b.fibers) if (!f.done) { any_live = true; if (f.wait != W_BARRIER) { at_barrier = false; break; } }
            if (any_live && at_barrier) { for (Fiber &f : b.fibers) if (!f.done) f.wait = W_NONE; continue; }
            if (!ran) {
                // a lane exited while its warp waits for it: the collective completes among the lanes that are left
                bool progressed = false;
                for (unsigned t = 0; t < nthreads; ++t) {
                    Fiber &f = b.fibers[t];
                    if (!f.done && f.wait == W_WARP && !f.resolved && try_resolve(b, t >> 5, f.mask, f.op)) progressed = true;
                }
                if (!progressed) {
                    unsigned nb = 0, nw = 0;
                    for (const Fiber &f : b.fibers) if (!f.done) { nb += f.wait == W_BARRIER; nw += f.wait == W_WARP; }
                    std::fprintf(stderr, "emu: deadlock in block %u: %u fibers at __syncthreads, %u at warp collectives\n", bx, nb, nw);
                    std::abort();
                }
            }
        }
    }
    g_block = nullptr; g_cur = nullptr;
}
}  // namespace emu
