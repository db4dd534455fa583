// TEST INFRASTRUCTURE (oracle/_ref only).  A minimal, FUNCTIONAL re-implementation of the part of EnTT's public interface
// the reference's sequential stepper uses, written from scratch from EnTT's documented behaviour (EnTT 3.15 is the
// pinned dependency, reference conanfile.py:77; it is neither installed nor vendored here and there is no network):
//   * entity = 20-bit index | 12-bit version; destroyed identifiers are recycled most-recently-destroyed first;
//   * one pool per component type: sparse set + packed array, "swap and pop" removal, iteration newest first;
//   * a multi-component view iterates its smallest pool and filters; each() skips empty (tag) types;
//   * signals on_construct (after), on_update (after), on_destroy (before the removal), listeners in reverse order.
// Not EnTT: no groups, no sorting, no runtime views, no meta reflection beyond the names some headers mention.
#pragma once
#include <algorithm>
#include <atomic>
#include <mutex>
#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <deque>
#include <functional>
#include <iterator>
#include <memory>
#include <new>
#include <string_view>
#include <tuple>
#include <type_traits>
#include <unordered_map>
#include <utility>
#include <vector>

namespace entt {

using id_type = std::uint32_t;

// ------------------------------------------------------------------------------------------------ type lists, tags
template<typename... T> struct type_list { static constexpr auto size = sizeof...(T); };
template<typename... T> struct exclude_t : type_list<T...> { explicit constexpr exclude_t() {} };
template<typename... T> struct get_t : type_list<T...> { explicit constexpr get_t() {} };
template<typename... T> inline constexpr exclude_t<T...> exclude{};
template<typename... T> inline constexpr get_t<T...> get{};
template<auto V> struct connect_arg_t { explicit connect_arg_t() = default; };
template<auto V> inline constexpr connect_arg_t<V> connect_arg{};
struct identity { template<typename T> constexpr T &&operator()(T &&v) const noexcept { return std::forward<T>(v); } };
template<typename T> struct type_identity { using type = T; };

// ------------------------------------------------------------------------------------------------ hashed strings, type info
class hashed_string {
    id_type h; const char *s;
public:
    static constexpr id_type value(const char *str) noexcept { id_type v = 2166136261u; while (*str) v = (v ^ static_cast<id_type>(*str++)) * 16777619u; return v; }
    static constexpr id_type value(const char *str, std::size_t n) noexcept { id_type v = 2166136261u; for (std::size_t i = 0; i < n; ++i) v = (v ^ static_cast<id_type>(str[i])) * 16777619u; return v; }
    constexpr hashed_string(const char *str) noexcept : h{value(str)}, s{str} {}
    constexpr hashed_string(const char *str, std::size_t n) noexcept : h{value(str, n)}, s{str} {}
    constexpr id_type value() const noexcept { return h; }
    constexpr operator id_type() const noexcept { return h; }
    constexpr const char *data() const noexcept { return s; }
};
inline namespace literals { constexpr hashed_string operator"" _hs(const char *str, std::size_t n) noexcept { return hashed_string{str, n}; } }

namespace internal {
inline id_type next_type_index() { static std::atomic<id_type> v{0}; return v++; }
inline std::uint64_t next_registry_uid() { static std::atomic<std::uint64_t> v{0}; return ++v; }
template<typename T> constexpr std::string_view pretty() { return __PRETTY_FUNCTION__; }
}
template<typename T> struct type_index { static id_type value() noexcept { static const id_type v = internal::next_type_index(); return v; } };
template<typename T> struct type_hash { static constexpr id_type value() noexcept { constexpr auto n = internal::pretty<T>(); return hashed_string::value(n.data(), n.size()); } };
template<typename T> struct type_name { static constexpr std::string_view value() noexcept { return internal::pretty<T>(); } };
struct type_info {
    id_type seq, id; std::string_view nm;
    id_type index() const noexcept { return seq; }
    id_type hash() const noexcept { return id; }
    std::string_view name() const noexcept { return nm; }
    bool operator==(const type_info &o) const noexcept { return id == o.id; }
    bool operator!=(const type_info &o) const noexcept { return id != o.id; }
};
template<typename T> const type_info &type_id() noexcept {
    using U = std::remove_cv_t<std::remove_reference_t<T>>;
    static const type_info info{type_index<U>::value(), type_hash<U>::value(), type_name<U>::value()};
    return info;
}
template<typename T> const type_info &type_id(T &&) noexcept { return type_id<std::remove_cv_t<std::remove_reference_t<T>>>(); }

// ------------------------------------------------------------------------------------------------ entities
enum class entity : std::uint32_t {};
struct entt_traits {
    using entity_type = std::uint32_t; using version_type = std::uint16_t;
    static constexpr entity_type entity_mask = 0xFFFFF, version_mask = 0xFFF; static constexpr int shift = 20;
};
constexpr std::uint32_t to_integral(entity e) noexcept { return static_cast<std::uint32_t>(e); }
constexpr std::uint32_t to_entity(entity e) noexcept { return to_integral(e) & entt_traits::entity_mask; }
constexpr std::uint32_t to_version(entity e) noexcept { return (to_integral(e) >> entt_traits::shift) & entt_traits::version_mask; }
constexpr entity make_entity(std::uint32_t idx, std::uint32_t ver) noexcept { return entity{(idx & entt_traits::entity_mask) | ((ver & entt_traits::version_mask) << entt_traits::shift)}; }
struct null_t {
    constexpr operator entity() const noexcept { return entity{0xFFFFFFFFu}; }
    constexpr bool operator==(null_t) const noexcept { return true; }
    constexpr bool operator!=(null_t) const noexcept { return false; }
    constexpr bool operator==(entity e) const noexcept { return to_entity(e) == entt_traits::entity_mask; }
    constexpr bool operator!=(entity e) const noexcept { return !(*this == e); }
};
constexpr bool operator==(entity e, null_t n) noexcept { return n == e; }
constexpr bool operator!=(entity e, null_t n) noexcept { return n != e; }
struct tombstone_t {
    constexpr operator entity() const noexcept { return entity{0xFFFFFFFFu}; }
    constexpr bool operator==(entity e) const noexcept { return to_version(e) == entt_traits::version_mask; }
    constexpr bool operator!=(entity e) const noexcept { return !(*this == e); }
};
constexpr bool operator==(entity e, tombstone_t n) noexcept { return n == e; }
constexpr bool operator!=(entity e, tombstone_t n) noexcept { return n != e; }
inline constexpr null_t null{};
inline constexpr tombstone_t tombstone{};

// ------------------------------------------------------------------------------------------------ signals
template<typename> class delegate;
template<typename Ret, typename... Args>
class delegate<Ret(Args...)> {
    using fn_t = Ret(const void *, Args...);
    fn_t *fn{}; const void *payload{};       // identity (for disconnect) = the trampoline instantiated for the connected callable + the instance
public:
    delegate() = default;
    template<auto C, typename... T> delegate(connect_arg_t<C>, T &&...inst) { connect<C>(std::forward<T>(inst)...); }
    // the callable may accept only the first N of the delegate's arguments (EnTT drops the trailing ones)
    template<auto C, typename... Head, std::size_t... I> static decltype(auto) call_n(std::index_sequence<I...>, std::tuple<Args &&...> &&t, Head &&...head) {
        return std::invoke(C, std::forward<Head>(head)..., std::forward<std::tuple_element_t<I, std::tuple<Args...>>>(std::get<I>(t))...);
    }
    template<auto C, std::size_t N, typename... Head> static constexpr bool callable_n(std::index_sequence<>) { return std::is_invocable_v<decltype(C), Head...>; }
    template<auto C, typename... Head, std::size_t... I> static constexpr bool invocable_with(std::index_sequence<I...>) {
        return std::is_invocable_v<decltype(C), Head..., std::tuple_element_t<I, std::tuple<Args...>>...>;
    }
    template<auto C, std::size_t N, typename... Head> static decltype(auto) call(std::tuple<Args &&...> &&t, Head &&...head) {
        if constexpr (invocable_with<C, Head...>(std::make_index_sequence<N>{})) return call_n<C>(std::make_index_sequence<N>{}, std::move(t), std::forward<Head>(head)...);
        else { static_assert(N > 0, "delegate: the callable does not accept these arguments"); return call<C, N - 1>(std::move(t), std::forward<Head>(head)...); }
    }
    template<auto C> void connect() noexcept {
        payload = nullptr;
        fn = [](const void *, Args... args) -> Ret { return Ret(call<C, sizeof...(Args)>(std::forward_as_tuple(std::forward<Args>(args)...))); };
    }
    template<auto C, typename T> void connect(T &inst) noexcept {
        payload = &inst;
        fn = [](const void *p, Args... args) -> Ret { return Ret(call<C, sizeof...(Args)>(std::forward_as_tuple(std::forward<Args>(args)...), *const_cast<T *>(static_cast<const T *>(p)))); };
    }
    template<auto C, typename T> void connect(T *inst) noexcept {
        payload = inst;
        fn = [](const void *p, Args... args) -> Ret { return Ret(call<C, sizeof...(Args)>(std::forward_as_tuple(std::forward<Args>(args)...), const_cast<T *>(static_cast<const T *>(p)))); };
    }
    void reset() noexcept { fn = nullptr; payload = nullptr; }
    const void *data() const noexcept { return payload; }
    Ret operator()(Args... args) const { return fn(payload, std::forward<Args>(args)...); }
    explicit operator bool() const noexcept { return fn != nullptr; }
    bool operator==(const delegate &o) const noexcept { return fn == o.fn && payload == o.payload; }
    bool operator!=(const delegate &o) const noexcept { return !(*this == o); }
};
template<typename> class sigh;
template<typename> class sink;
template<typename Ret, typename... Args>
class sigh<Ret(Args...)> {
    friend class sink<sigh<Ret(Args...)>>;
    std::vector<delegate<Ret(Args...)>> calls;
public:
    using sink_type = sink<sigh<Ret(Args...)>>;
    std::size_t size() const noexcept { return calls.size(); }
    bool empty() const noexcept { return calls.empty(); }
    void publish(Args... args) const { for (auto pos = calls.size(); pos; --pos) calls[pos - 1](args...); }      // listeners may disconnect themselves
    template<typename F> void collect(F f, Args... args) const { for (auto pos = calls.size(); pos; --pos) f(calls[pos - 1](args...)); }
};
class connection {
    std::function<void()> undo;
public:
    connection() = default;
    explicit connection(std::function<void()> f) : undo{std::move(f)} {}
    explicit operator bool() const noexcept { return static_cast<bool>(undo); }
    void release() { if (undo) { undo(); undo = nullptr; } }
};
struct scoped_connection {
    scoped_connection() = default;
    scoped_connection(const connection &c) : conn{c} {}
    scoped_connection(const scoped_connection &) = delete;
    scoped_connection(scoped_connection &&o) noexcept : conn{std::exchange(o.conn, {})} {}
    ~scoped_connection() { conn.release(); }
    scoped_connection &operator=(const scoped_connection &) = delete;
    scoped_connection &operator=(scoped_connection &&o) noexcept { conn.release(); conn = std::exchange(o.conn, {}); return *this; }
    scoped_connection &operator=(connection c) { conn.release(); conn = std::move(c); return *this; }
    explicit operator bool() const noexcept { return static_cast<bool>(conn); }
    void release() { conn.release(); }
private:
    connection conn;
};
template<typename Ret, typename... Args>
class sink<sigh<Ret(Args...)>> {
    using signal_type = sigh<Ret(Args...)>;
    using delegate_type = delegate<Ret(Args...)>;
    signal_type *sig;
    void drop(const delegate_type &d) { auto &c = sig->calls; c.erase(std::remove(c.begin(), c.end(), d), c.end()); }
public:
    sink(signal_type &s) noexcept : sig{&s} {}
    bool empty() const noexcept { return sig->calls.empty(); }
    template<auto C, typename... T> connection connect(T &&...inst) {
        delegate_type d; d.template connect<C>(std::forward<T>(inst)...);
        drop(d);                                       // connecting twice keeps one
        sig->calls.push_back(d);
        signal_type *s = sig;
        return connection{[s, d]() { auto &c = s->calls; c.erase(std::remove(c.begin(), c.end(), d), c.end()); }};
    }
    template<auto C, typename... T> void disconnect(T &&...inst) { delegate_type d; d.template connect<C>(std::forward<T>(inst)...); drop(d); }
    template<typename T> void disconnect(T &inst) { const void *p = &inst; auto &c = sig->calls; c.erase(std::remove_if(c.begin(), c.end(), [p](const delegate_type &d) { return d.data() == p; }), c.end()); }
    template<typename T> void disconnect(T *inst) { const void *p = inst; auto &c = sig->calls; c.erase(std::remove_if(c.begin(), c.end(), [p](const delegate_type &d) { return d.data() == p; }), c.end()); }
    void disconnect() { sig->calls.clear(); }
};
template<typename Ret, typename... Args> sink(sigh<Ret(Args...)> &) -> sink<sigh<Ret(Args...)>>;

// ------------------------------------------------------------------------------------------------ sparse set
class sparse_set {
protected:
    static constexpr std::uint32_t npos = 0xFFFFFFFFu;
    static constexpr std::size_t page_bits = 12;      // EnTT: ENTT_SPARSE_PAGE = 4096 entries; a set of a few entities touches a few pages
    struct sparse_pages {
        std::vector<std::vector<std::uint32_t>> pages;
        std::uint32_t get(std::size_t idx) const noexcept {
            const auto pg = idx >> page_bits;
            return pg < pages.size() && !pages[pg].empty() ? pages[pg][idx & ((std::size_t{1} << page_bits) - 1)] : npos;
        }
        std::uint32_t &operator[](std::size_t idx) {
            const auto pg = idx >> page_bits;
            if (pg >= pages.size()) pages.resize(pg + 1);
            if (pages[pg].empty()) pages[pg].assign(std::size_t{1} << page_bits, npos);
            return pages[pg][idx & ((std::size_t{1} << page_bits) - 1)];
        }
        void reset() noexcept { pages.clear(); }
        void swap(sparse_pages &o) noexcept { pages.swap(o.pages); }
    };
    std::vector<entity> packed;
    sparse_pages sparse;
    virtual void swap_and_pop(std::size_t pos) {        // the last element takes the place of the removed one
        const entity last = packed.back();
        sparse[to_entity(packed[pos])] = npos;
        if (pos != packed.size() - 1) { packed[pos] = last; sparse[to_entity(last)] = static_cast<std::uint32_t>(pos); }
        packed.pop_back();
    }
    virtual void on_clear() {}
    void append(entity e) {
        sparse[to_entity(e)] = static_cast<std::uint32_t>(packed.size());
        packed.push_back(e);
    }
public:
    using entity_type = entity;
    using size_type = std::size_t;
    using iterator = std::vector<entity>::const_reverse_iterator;        // newest first
    using const_iterator = iterator;
    using reverse_iterator = std::vector<entity>::const_iterator;
    sparse_set() = default;
    sparse_set(const sparse_set &) = default;
    sparse_set(sparse_set &&) = default;
    sparse_set &operator=(const sparse_set &) = default;
    sparse_set &operator=(sparse_set &&) = default;
    virtual ~sparse_set() = default;
    size_type size() const noexcept { return packed.size(); }
    bool empty() const noexcept { return packed.empty(); }
    const entity *data() const noexcept { return packed.data(); }
    void reserve(size_type n) { packed.reserve(n); }
    size_type capacity() const noexcept { return packed.capacity(); }
    iterator begin() const noexcept { return packed.rbegin(); }
    iterator end() const noexcept { return packed.rend(); }
    iterator cbegin() const noexcept { return begin(); }
    iterator cend() const noexcept { return end(); }
    reverse_iterator rbegin() const noexcept { return packed.begin(); }
    reverse_iterator rend() const noexcept { return packed.end(); }
    bool contains(entity e) const noexcept { const auto pos = sparse.get(to_entity(e)); return pos != npos && packed[pos] == e; }
    size_type index(entity e) const noexcept { return sparse.get(to_entity(e)); }
    iterator find(entity e) const noexcept { return contains(e) ? iterator{packed.begin() + index(e) + 1} : end(); }
    entity at(size_type pos) const noexcept { return pos < packed.size() ? packed[pos] : entity{null}; }
    entity operator[](size_type pos) const noexcept { return packed[pos]; }
    iterator push(entity e) { append(e); return begin(); }
    template<typename It> iterator push(It first, It last) { for (; first != last; ++first) append(*first); return begin(); }
    iterator emplace(entity e) { return push(e); }
    template<typename It> void insert(It first, It last) { push(first, last); }
    void erase(entity e) { swap_and_pop(index(e)); }
    template<typename It> void erase(It first, It last) { for (; first != last; ++first) erase(*first); }
    bool remove(entity e) { return contains(e) ? (erase(e), true) : false; }
    template<typename It> size_type remove(It first, It last) { size_type n = 0; for (; first != last; ++first) n += remove(*first); return n; }
    void clear() { on_clear(); packed.clear(); sparse.reset(); }
    void swap(sparse_set &o) noexcept { packed.swap(o.packed); sparse.swap(o.sparse); }
};

// ------------------------------------------------------------------------------------------------ paged component array
// EnTT keeps components in pages of ENTT_PACKED_PAGE (1024) elements so that references survive appends; same here.
template<typename T>
class paged_vector {
    static constexpr std::size_t page_bits = 10, page_size = std::size_t{1} << page_bits;
    std::vector<T *> pages;
    std::size_t count{};
    T *slot(std::size_t i) const noexcept { return pages[i >> page_bits] + (i & (page_size - 1)); }
    void release() noexcept { clear(); for (T *p : pages) ::operator delete(p, std::align_val_t{alignof(T)}); pages.clear(); }
public:
    paged_vector() = default;
    paged_vector(const paged_vector &o) { for (std::size_t i = 0; i < o.count; ++i) emplace_back(o[i]); }
    paged_vector(paged_vector &&o) noexcept : pages(std::move(o.pages)), count(o.count) { o.pages.clear(); o.count = 0; }
    paged_vector &operator=(paged_vector o) noexcept { pages.swap(o.pages); std::swap(count, o.count); return *this; }
    ~paged_vector() { release(); }
    std::size_t size() const noexcept { return count; }
    T &operator[](std::size_t i) noexcept { return *slot(i); }
    const T &operator[](std::size_t i) const noexcept { return *slot(i); }
    T &back() noexcept { return *slot(count - 1); }
    template<typename... A> T &emplace_back(A &&...args) {
        if ((count >> page_bits) == pages.size()) pages.push_back(static_cast<T *>(::operator new(sizeof(T) * page_size, std::align_val_t{alignof(T)})));
        T *p = ::new (static_cast<void *>(slot(count))) T(std::forward<A>(args)...);
        ++count;
        return *p;
    }
    void push_back(T &&v) { emplace_back(std::move(v)); }
    void pop_back() noexcept { slot(--count)->~T(); }
    void clear() noexcept { while (count) pop_back(); }
};

// ------------------------------------------------------------------------------------------------ component pools
template<typename Registry>
class basic_pool : public sparse_set {           // what the registry needs from a pool without knowing its type
public:
    Registry *owner{};
    sigh<void(Registry &, entity)> construction, update, destruction;
    auto on_construct() noexcept { return sink{construction}; }
    auto on_update() noexcept { return sink{update}; }
    auto on_destroy() noexcept { return sink{destruction}; }
    virtual bool remove_from(entity e) = 0;       // with the destruction signal
    virtual void clear_all() = 0;
};
template<typename T, typename Registry>
class storage_impl final : public basic_pool<Registry> {
    static constexpr bool has_payload = !std::is_empty_v<T>;
    using base = basic_pool<Registry>;
    paged_vector<T> payload;                      // references stay valid while elements are appended
    void swap_and_pop(std::size_t pos) override {
        if constexpr (has_payload) { if (pos != payload.size() - 1) payload[pos] = std::move(payload.back()); payload.pop_back(); }
        sparse_set::swap_and_pop(pos);
    }
    void on_clear() override { payload.clear(); }
public:
    using value_type = T;
    using entity_type = entity;
    template<typename... A> decltype(auto) emplace(entity e, A &&...args) {
        if constexpr (has_payload) {
            if constexpr (std::is_aggregate_v<T> && (sizeof...(A) != 0 || !std::is_default_constructible_v<T>)) payload.push_back(T{std::forward<A>(args)...});
            else payload.emplace_back(std::forward<A>(args)...);
        }
        this->append(e);
        this->construction.publish(*this->owner, e);
        if constexpr (has_payload) return static_cast<T &>(payload[this->index(e)]);
    }
    template<typename It> void insert(It first, It last, const T &value = {}) { for (; first != last; ++first) emplace(*first, value); }
    decltype(auto) get(entity e) noexcept { if constexpr (has_payload) return static_cast<T &>(payload[this->index(e)]); }
    decltype(auto) get(entity e) const noexcept { if constexpr (has_payload) return static_cast<const T &>(payload[this->index(e)]); }
    auto get_as_tuple(entity e) noexcept { if constexpr (has_payload) return std::forward_as_tuple(get(e)); else return std::tuple<>{}; }
    auto get_as_tuple(entity e) const noexcept { if constexpr (has_payload) return std::forward_as_tuple(get(e)); else return std::tuple<>{}; }
    template<typename... F> decltype(auto) patch(entity e, F &&...func) {
        if constexpr (has_payload) { auto &v = payload[this->index(e)]; (std::forward<F>(func)(v), ...); this->update.publish(*this->owner, e); return static_cast<T &>(v); }
        else { this->update.publish(*this->owner, e); }
    }
    void erase(entity e) { this->destruction.publish(*this->owner, e); sparse_set::erase(e); }
    template<typename It> void erase(It first, It last) { for (; first != last; ++first) erase(*first); }
    bool remove(entity e) { return this->contains(e) ? (erase(e), true) : false; }
    template<typename It> std::size_t remove(It first, It last) { std::size_t n = 0; for (; first != last; ++first) n += remove(*first); return n; }
    bool remove_from(entity e) override { return remove(e); }
    void clear_all() override { while (!this->empty()) erase(this->packed.back()); }
    void clear() { clear_all(); }
    // (entity, component) pairs, newest first
    struct iterable {
        storage_impl *s;
        struct iterator {
            storage_impl *s; std::size_t left;
            using value_type = decltype(std::tuple_cat(std::make_tuple(entity{}), std::declval<storage_impl &>().get_as_tuple(entity{})));
            using difference_type = std::ptrdiff_t; using pointer = void; using reference = value_type; using iterator_category = std::input_iterator_tag;
            value_type operator*() const { const entity e = s->packed[left - 1]; return std::tuple_cat(std::make_tuple(e), s->get_as_tuple(e)); }
            iterator &operator++() { --left; return *this; }
            bool operator==(const iterator &o) const { return left == o.left; }
            bool operator!=(const iterator &o) const { return left != o.left; }
        };
        iterator begin() const { return {s, s->packed.size()}; }
        iterator end() const { return {s, 0}; }
    };
    iterable each() noexcept { return {this}; }
};

// ------------------------------------------------------------------------------------------------ views
template<typename, typename> class basic_view;
template<typename... Get, typename... Exclude>
class basic_view<get_t<Get...>, exclude_t<Exclude...>> {
    static_assert(sizeof...(Get) > 0);
    std::tuple<Get *...> pools;
    std::tuple<Exclude *...> filter;
    const sparse_set *lead{};
    template<typename T> static constexpr std::size_t index_of() {
        std::size_t i = 0; bool found = false;
        ((found || (std::is_same_v<std::remove_const_t<T>, typename Get::value_type> ? (found = true) : (++i, false))), ...);
        return i;
    }
    void pick_lead() noexcept {        // the smallest pool drives the iteration (the first of equals)
        lead = nullptr;
        std::apply([this](auto *...p) { ((lead = (lead == nullptr || p->size() < lead->size()) ? static_cast<const sparse_set *>(p) : lead), ...); }, pools);
    }
    bool accept(entity e) const noexcept {
        return std::apply([e](auto *...p) { return (p->contains(e) && ...); }, pools) && std::apply([e](auto *...p) { return !(p->contains(e) || ...); }, filter);
    }
    auto tuple_of(entity e) const { return std::apply([e](auto *...p) { return std::tuple_cat(p->get_as_tuple(e)...); }, pools); }
public:
    using entity_type = entity;
    using size_type = std::size_t;
    basic_view() noexcept : pools{}, filter{} {}
    basic_view(std::tuple<Get *...> g, std::tuple<Exclude *...> x = {}) noexcept : pools{g}, filter{x} { pick_lead(); }
    basic_view(Get &...g, Exclude &...x) noexcept : pools{&g...}, filter{&x...} { pick_lead(); }
    class iterator {
        const basic_view *v; sparse_set::iterator it, last;
        void skip() { while (it != last && !v->accept(*it)) ++it; }
    public:
        using value_type = entity; using difference_type = std::ptrdiff_t; using pointer = const entity *; using reference = entity; using iterator_category = std::forward_iterator_tag;
        iterator() : v{}, it{}, last{} {}
        iterator(const basic_view *vw, sparse_set::iterator b, sparse_set::iterator e) : v{vw}, it{b}, last{e} { skip(); }
        entity operator*() const { return *it; }
        iterator &operator++() { ++it; skip(); return *this; }
        iterator operator++(int) { iterator c = *this; ++*this; return c; }
        bool operator==(const iterator &o) const { return it == o.it; }
        bool operator!=(const iterator &o) const { return it != o.it; }
    };
    iterator begin() const { return lead ? iterator{this, lead->begin(), lead->end()} : iterator{}; }
    iterator end() const { return lead ? iterator{this, lead->end(), lead->end()} : iterator{}; }
    size_type size_hint() const noexcept { return lead ? lead->size() : 0; }
    template<std::size_t N = sizeof...(Get) + sizeof...(Exclude), std::enable_if_t<N == 1, int> = 0> size_type size() const noexcept { return lead ? lead->size() : 0; }
    template<std::size_t N = sizeof...(Get) + sizeof...(Exclude), std::enable_if_t<N == 1, int> = 0> bool empty() const noexcept { return !lead || lead->empty(); }
    explicit operator bool() const noexcept { return lead != nullptr; }
    entity front() const { auto it = begin(); return it != end() ? *it : entity{null}; }
    entity back() const { entity last = null; for (auto e : *this) last = e; return last; }
    bool contains(entity e) const noexcept { return lead && accept(e); }
    iterator find(entity e) const { return contains(e) ? iterator{this, lead->find(e), lead->end()} : end(); }
    template<typename T> void use() noexcept { lead = std::get<index_of<T>()>(pools); }
    template<std::size_t I> void use() noexcept { lead = std::get<I>(pools); }
    template<typename T> auto &storage() const noexcept { return *std::get<index_of<T>()>(pools); }
    template<std::size_t I> auto &storage() const noexcept { return *std::get<I>(pools); }
    template<typename... T> decltype(auto) get(entity e) const {
        if constexpr (sizeof...(T) == 0) return tuple_of(e);
        else if constexpr (sizeof...(T) == 1) return (std::get<index_of<T>()>(pools)->get(e), ...);
        else return std::tuple_cat(std::get<index_of<T>()>(pools)->get_as_tuple(e)...);
    }
    template<std::size_t I, std::size_t... J> decltype(auto) get(entity e) const {
        if constexpr (sizeof...(J) == 0) return std::get<I>(pools)->get(e);
        else return std::tuple_cat(std::get<I>(pools)->get_as_tuple(e), std::get<J>(pools)->get_as_tuple(e)...);
    }
    decltype(auto) operator[](entity e) const { return get<>(e); }
    template<typename F> void each(F func) const {        // func(entity, components...) or func(components...); empty types are skipped
        for (const entity e : *this) {
            auto comps = tuple_of(e);
            if constexpr (can_apply_with_entity<F, decltype(comps)>::value) std::apply(func, std::tuple_cat(std::make_tuple(e), comps));
            else std::apply(func, comps);
        }
    }
    struct iterable {                 // holds its own copy of the view: `for (auto [e, ...] : registry.view<...>().each())`
        basic_view vw;
        struct iterator {
            const basic_view *v; typename basic_view::iterator it;
            using value_type = decltype(std::tuple_cat(std::make_tuple(entity{}), std::declval<const basic_view &>().tuple_of(entity{})));
            using difference_type = std::ptrdiff_t; using pointer = void; using reference = value_type; using iterator_category = std::input_iterator_tag;
            value_type operator*() const { const entity e = *it; return std::tuple_cat(std::make_tuple(e), v->tuple_of(e)); }
            iterator &operator++() { ++it; return *this; }
            bool operator==(const iterator &o) const { return it == o.it; }
            bool operator!=(const iterator &o) const { return it != o.it; }
        };
        iterator begin() const { return {&vw, vw.begin()}; }
        iterator end() const { return {&vw, vw.end()}; }
    };
    iterable each() const noexcept { return {*this}; }
    template<typename... OG, typename... OE>
    auto operator|(const basic_view<get_t<OG...>, exclude_t<OE...>> &o) const noexcept {
        return basic_view<get_t<Get..., OG...>, exclude_t<Exclude..., OE...>>{std::tuple_cat(pools, o.raw_pools()), std::tuple_cat(filter, o.raw_filter())};
    }
    const auto &raw_pools() const noexcept { return pools; }
    const auto &raw_filter() const noexcept { return filter; }
private:
    template<typename F, typename Tuple> struct can_apply_with_entity;
    template<typename F, typename... C> struct can_apply_with_entity<F, std::tuple<C...>> : std::is_invocable<F &, entity, C...> {};
};

// ------------------------------------------------------------------------------------------------ registry
class registry {
    using pool_base = basic_pool<registry>;
public:
    using entity_type = entity;
    using size_type = std::size_t;
    using allocator_type = std::allocator<entity>;
    template<typename T> using storage_for_type = std::conditional_t<std::is_const_v<T>, const storage_impl<std::remove_const_t<T>, registry>, storage_impl<std::remove_const_t<T>, registry>>;

    class context {
        std::unordered_map<id_type, std::shared_ptr<void>> vars;
    public:
        context() = default;
        explicit context(const allocator_type &) {}
        template<typename T, typename... A> T &emplace(A &&...args) { return emplace_as<T>(type_hash<T>::value(), std::forward<A>(args)...); }
        template<typename T, typename... A> T &emplace_as(id_type id, A &&...args) {
            auto it = vars.find(id);
            if (it == vars.end()) { std::shared_ptr<T> p; if constexpr (std::is_aggregate_v<T>) p.reset(new T{std::forward<A>(args)...}); else p.reset(new T(std::forward<A>(args)...)); it = vars.emplace(id, std::static_pointer_cast<void>(p)).first; }
            return *static_cast<T *>(it->second.get());
        }
        template<typename T> T &insert_or_assign(T &&v) { using U = std::remove_cv_t<std::remove_reference_t<T>>; vars[type_hash<U>::value()] = std::static_pointer_cast<void>(std::make_shared<U>(std::forward<T>(v))); return *static_cast<U *>(vars[type_hash<U>::value()].get()); }
        template<typename T> bool erase(id_type id = type_hash<T>::value()) { return vars.erase(id) != 0; }
        template<typename T> T &get(id_type id = type_hash<std::remove_const_t<T>>::value()) { return *static_cast<T *>(vars.at(id).get()); }
        template<typename T> const T &get(id_type id = type_hash<std::remove_const_t<T>>::value()) const { return *static_cast<const T *>(vars.at(id).get()); }
        template<typename T> T *find(id_type id = type_hash<std::remove_const_t<T>>::value()) { auto it = vars.find(id); return it == vars.end() ? nullptr : static_cast<T *>(it->second.get()); }
        template<typename T> const T *find(id_type id = type_hash<std::remove_const_t<T>>::value()) const { auto it = vars.find(id); return it == vars.end() ? nullptr : static_cast<const T *>(it->second.get()); }
        template<typename T> bool contains(id_type id = type_hash<std::remove_const_t<T>>::value()) const { return vars.count(id) != 0; }
    };

    registry() = default;
    registry(const registry &) = delete;
    registry &operator=(const registry &) = delete;

    // ---- entities.  Alive identifiers occupy [0, alive) of `ids`; the most recently destroyed one is recycled first.
    entity create() {
        if (alive < ids.size()) return ids[alive++];
        const entity e = make_entity(static_cast<std::uint32_t>(ids.size()), 0);
        where.push_back(static_cast<std::uint32_t>(ids.size())); ids.push_back(e); ++alive;
        return e;
    }
    template<typename It> void create(It first, It last) { for (; first != last; ++first) *first = create(); }
    bool valid(entity e) const noexcept { const auto idx = to_entity(e); return idx < where.size() && where[idx] < alive && ids[where[idx]] == e; }
    entity current(entity e) const noexcept { return ids[where[to_entity(e)]]; }
    void destroy(entity e) {
        for (auto pos = order.size(); pos; --pos) order[pos - 1]->remove_from(e);      // pools in reverse order of creation
        release(e);
    }
    template<typename It> void destroy(It first, It last) { std::vector<entity> tmp(first, last); for (entity e : tmp) destroy(e); }
    void release(entity e) {
        const auto idx = to_entity(e);
        const auto pos = where[idx], lastpos = static_cast<std::uint32_t>(alive - 1);
        const entity moved = ids[lastpos];
        ids[pos] = moved; where[to_entity(moved)] = pos;
        ids[lastpos] = make_entity(idx, to_version(e) + 1); where[idx] = lastpos;
        --alive;
    }
    size_type alive_count() const noexcept { return alive; }

    // ---- pools
    template<typename T> storage_for_type<std::remove_const_t<T>> &storage() { return assure<std::remove_const_t<T>>(); }
    template<typename T> const storage_for_type<std::remove_const_t<T>> *storage() const {
        using pool_type = storage_impl<std::remove_const_t<T>, registry>;
        static thread_local std::uint64_t cached_owner = 0;
        static thread_local const pool_type *cached = nullptr;
        if (cached_owner == uid) return cached;
        std::lock_guard<std::mutex> lock(pools_mutex);
        auto it = pools.find(type_hash<std::remove_const_t<T>>::value());
        if (it == pools.end()) return nullptr;
        cached = static_cast<const pool_type *>(it->second.get()); cached_owner = uid;
        return cached;
    }
    template<typename T, typename... A> decltype(auto) emplace(entity e, A &&...args) { return assure<T>().emplace(e, std::forward<A>(args)...); }
    template<typename T, typename It> void insert(It first, It last, const T &value = {}) { assure<T>().insert(first, last, value); }
    template<typename T, typename... A> decltype(auto) emplace_or_replace(entity e, A &&...args) {
        auto &p = assure<T>();
        if (p.contains(e)) return replace<T>(e, std::forward<A>(args)...);
        return p.emplace(e, std::forward<A>(args)...);
    }
    template<typename T, typename... F> decltype(auto) patch(entity e, F &&...func) { return assure<T>().patch(e, std::forward<F>(func)...); }
    template<typename T, typename... A> decltype(auto) replace(entity e, A &&...args) {
        if constexpr (std::is_empty_v<T>) return assure<T>().patch(e);
        else return assure<T>().patch(e, [&args...](auto &v) { if constexpr (std::is_aggregate_v<T>) v = T{std::forward<A>(args)...}; else v = T(std::forward<A>(args)...); });
    }
    template<typename T, typename... Other> size_type remove(entity e) { return (assure<T>().remove(e) + ... + assure<Other>().remove(e)); }
    template<typename T, typename... Other, typename It> size_type remove(It first, It last) {
        std::vector<entity> tmp(first, last); size_type n = 0;
        for (entity e : tmp) n += remove<T, Other...>(e);
        return n;
    }
    template<typename T, typename... Other> void erase(entity e) { (assure<T>().erase(e), (assure<Other>().erase(e), ...)); }
    template<typename T, typename... Other, typename It> void erase(It first, It last) { std::vector<entity> tmp(first, last); for (entity e : tmp) erase<T, Other...>(e); }
    template<typename... T> bool all_of(entity e) const noexcept { return (has<T>(e) && ...); }
    template<typename... T> bool any_of(entity e) const noexcept { return (has<T>(e) || ...); }
    template<typename... T> decltype(auto) get(entity e) {
        if constexpr (sizeof...(T) == 1) return (assure<std::remove_const_t<T>>().get(e), ...);
        else return std::forward_as_tuple(assure<std::remove_const_t<T>>().get(e)...);
    }
    template<typename... T> decltype(auto) get(entity e) const {
        if constexpr (sizeof...(T) == 1) return (storage<T>()->get(e), ...);
        else return std::forward_as_tuple(storage<T>()->get(e)...);
    }
    template<typename T, typename... A> decltype(auto) get_or_emplace(entity e, A &&...args) { auto &p = assure<T>(); return p.contains(e) ? p.get(e) : p.emplace(e, std::forward<A>(args)...); }
    template<typename... T> auto try_get(entity e) {
        if constexpr (sizeof...(T) == 1) { auto &p = (assure<std::remove_const_t<T>>(), ...); return p.contains(e) ? std::addressof(p.get(e)) : nullptr; }
        else return std::make_tuple(try_get<T>(e)...);
    }
    template<typename... T> auto try_get(entity e) const {
        if constexpr (sizeof...(T) == 1) { auto *p = (storage<T>(), ...); return (p && p->contains(e)) ? std::addressof(p->get(e)) : nullptr; }
        else return std::make_tuple(try_get<T>(e)...);
    }
    template<typename... T> void clear() {
        if constexpr (sizeof...(T) == 0) { for (auto pos = order.size(); pos; --pos) order[pos - 1]->clear_all(); while (alive) release(ids[alive - 1]); }
        else (assure<T>().clear_all(), ...);
    }
    bool orphan(entity e) const noexcept { for (auto *p : order) if (p->contains(e)) return false; return true; }
    template<typename T> auto on_construct() { return assure<T>().on_construct(); }
    template<typename T> auto on_update() { return assure<T>().on_update(); }
    template<typename T> auto on_destroy() { return assure<T>().on_destroy(); }

    template<typename T, typename... Other, typename... E>
    basic_view<get_t<storage_for_type<T>, storage_for_type<Other>...>, exclude_t<storage_for_type<E>...>> view(exclude_t<E...> = exclude_t<>{}) {
        return {assure<std::remove_const_t<T>>(), assure<std::remove_const_t<Other>>()..., assure<std::remove_const_t<E>>()...};
    }
    template<typename T, typename... Other, typename... E>
    basic_view<get_t<storage_for_type<const T>, storage_for_type<const Other>...>, exclude_t<storage_for_type<const E>...>> view(exclude_t<E...> = exclude_t<>{}) const {
        auto &self = const_cast<registry &>(*this);
        return {self.assure<std::remove_const_t<T>>(), self.assure<std::remove_const_t<Other>>()..., self.assure<std::remove_const_t<E>>()...};
    }
    context &ctx() noexcept { return vars; }
    const context &ctx() const noexcept { return vars; }

private:
    template<typename T> bool has(entity e) const noexcept { auto *p = storage<T>(); return p && p->contains(e); }
    template<typename T> storage_impl<T, registry> &assure() {
        static_assert(!std::is_const_v<T>);
        // Pools are created on first use and never removed.  The sequential_multithreaded mode calls registry.view<...>()
        // from worker threads: the lookup goes through a per-thread cache keyed by the registry's unique id, creation and
        // map access sit behind a mutex (EnTT itself leaves this to the caller).
        static thread_local std::uint64_t cached_owner = 0;
        static thread_local storage_impl<T, registry> *cached = nullptr;
        if (cached_owner == uid) return *cached;
        std::lock_guard<std::mutex> lock(pools_mutex);
        const id_type id = type_hash<T>::value();
        auto it = pools.find(id);
        if (it == pools.end()) {
            auto p = std::make_unique<storage_impl<T, registry>>();
            p->owner = this;
            order.push_back(p.get());
            it = pools.emplace(id, std::move(p)).first;
        }
        cached = static_cast<storage_impl<T, registry> *>(it->second.get()); cached_owner = uid;
        return *cached;
    }
    const std::uint64_t uid = internal::next_registry_uid();
    mutable std::mutex pools_mutex;
    std::vector<entity> ids; std::vector<std::uint32_t> where; size_type alive{};
    std::unordered_map<id_type, std::unique_ptr<pool_base>> pools;
    std::vector<pool_base *> order;
    context vars;
};
using basic_registry = registry;
template<typename T> using storage = storage_impl<T, registry>;
template<typename T> using storage_for_t = registry::storage_for_type<T>;
template<typename Get, typename Exclude = exclude_t<>> using view = basic_view<Get, Exclude>;
struct as_view { registry &reg; template<typename G, typename E> operator basic_view<G, E>() const; };

template<typename K, typename V, typename H = std::hash<K>, typename E = std::equal_to<K>> using dense_map = std::unordered_map<K, V, H, E>;

// ------------------------------------------------------------------------------------------------ any
class any {
    std::shared_ptr<void> p; const type_info *ti{};
public:
    any() = default;
    template<typename T, typename... A> explicit any(std::in_place_type_t<T>, A &&...args) : ti{&type_id<T>()} {
        if constexpr (std::is_aggregate_v<T>) p = std::shared_ptr<T>(new T{std::forward<A>(args)...}); else p = std::make_shared<T>(std::forward<A>(args)...);
    }
    template<typename T, typename = std::enable_if_t<!std::is_same_v<std::decay_t<T>, any>>> any(T &&v) : p{std::make_shared<std::decay_t<T>>(std::forward<T>(v))}, ti{&type_id<std::decay_t<T>>()} {}
    explicit operator bool() const noexcept { return static_cast<bool>(p); }
    const type_info &type() const noexcept { return ti ? *ti : type_id<void>(); }
    void *data() noexcept { return p.get(); }
    const void *data() const noexcept { return p.get(); }
    void reset() { p.reset(); ti = nullptr; }
};
template<typename T> T *any_cast(any *a) noexcept { return (a && a->type() == type_id<T>()) ? static_cast<T *>(a->data()) : nullptr; }
template<typename T> const T *any_cast(const any *a) noexcept { return (a && a->type() == type_id<T>()) ? static_cast<const T *>(a->data()) : nullptr; }

// ------------------------------------------------------------------------------------------------ meta: NAMES ONLY.
// The reference's replication / networking code uses EnTT's runtime reflection; the sequential stepper never reaches
// it.  Every function here aborts with a message: the headers parse, the library links, nothing is silently wrong.
namespace internal { [[noreturn]] inline void no_meta() { std::fputs("entt_lite: entt::meta is not implemented (replication / networking code reached)\n", stderr); std::abort(); } }
#define ENTT_LITE_NO_META { ::entt::internal::no_meta(); }
class meta_any; class meta_type; class meta_data; class meta_handle; class meta_sequence_container; class meta_associative_container;
template<typename T> struct meta_range {
    struct iterator { std::pair<id_type, T> operator*() const ENTT_LITE_NO_META iterator &operator++() ENTT_LITE_NO_META bool operator!=(const iterator &) const ENTT_LITE_NO_META bool operator==(const iterator &) const ENTT_LITE_NO_META };
    iterator begin() const ENTT_LITE_NO_META iterator end() const ENTT_LITE_NO_META
};
class meta_type {
public:
    meta_type() = default;
    id_type id() const; const type_info &info() const;
    meta_range<meta_data> data() const; meta_data data(id_type) const;
    bool is_sequence_container() const; bool is_associative_container() const; bool is_class() const; bool is_arithmetic() const;
    explicit operator bool() const; bool operator==(const meta_type &) const; bool operator!=(const meta_type &) const;
    meta_any construct() const; template<typename... A> meta_any construct(A &&...) const;
    template<typename... A> meta_any invoke(id_type, meta_handle, A &&...) const;
    meta_any from_void(void *) const; meta_any from_void(const void *) const;
};
class meta_handle { public: meta_handle() = default; template<typename T> meta_handle(T &) ENTT_LITE_NO_META meta_handle(meta_any &); meta_handle(const meta_any &); explicit operator bool() const; meta_any *operator->(); };
class meta_sequence_container {
public:
    struct iterator { meta_any operator*() const; iterator &operator++(); bool operator!=(const iterator &) const; bool operator==(const iterator &) const; explicit operator bool() const; };
    meta_type value_type() const; std::size_t size() const; bool resize(std::size_t); bool clear(); bool reserve(std::size_t);
    iterator begin(); iterator end(); iterator insert(iterator, meta_any); iterator erase(iterator); meta_any operator[](std::size_t);
    explicit operator bool() const;
};
class meta_associative_container {
public:
    struct iterator { std::pair<meta_any, meta_any> operator*() const; iterator &operator++(); bool operator!=(const iterator &) const; bool operator==(const iterator &) const; explicit operator bool() const; };
    meta_type key_type() const; meta_type mapped_type() const; meta_type value_type() const; std::size_t size() const; bool clear();
    iterator begin(); iterator end(); bool insert(meta_any, meta_any); std::size_t erase(meta_any); iterator find(meta_any);
    explicit operator bool() const;
};
class meta_any {
public:
    meta_any() = default;
    template<typename T, typename = std::enable_if_t<!std::is_same_v<std::decay_t<T>, meta_any>>> meta_any(T &&) ENTT_LITE_NO_META
    template<typename T, typename... A> explicit meta_any(std::in_place_type_t<T>, A &&...) ENTT_LITE_NO_META
    meta_type type() const;
    template<typename T> T cast() const ENTT_LITE_NO_META template<typename T> T cast() ENTT_LITE_NO_META
    template<typename T> const T *try_cast() const ENTT_LITE_NO_META template<typename T> T *try_cast() ENTT_LITE_NO_META
    template<typename T> bool allow_cast() ENTT_LITE_NO_META
    meta_sequence_container as_sequence_container(); meta_sequence_container as_sequence_container() const;
    meta_associative_container as_associative_container(); meta_associative_container as_associative_container() const;
    meta_any as_ref(); meta_any as_ref() const;
    template<typename... A> meta_any invoke(id_type, A &&...) const ENTT_LITE_NO_META
    template<typename T> bool set(id_type, T &&) ENTT_LITE_NO_META meta_any get(id_type) const;
    void *data(); const void *data() const;
    explicit operator bool() const; bool operator==(const meta_any &) const; bool operator!=(const meta_any &) const;
    template<typename T> void assign(T &&) ENTT_LITE_NO_META void reset();
};
class meta_data {
public:
    meta_type type() const; meta_any get(meta_handle) const; template<typename T> bool set(meta_handle, T &&) const ENTT_LITE_NO_META
    bool is_const() const; bool is_static() const; explicit operator bool() const;
};
// out-of-line: the classes refer to each other by value
template<typename... A> meta_any meta_type::construct(A &&...) const ENTT_LITE_NO_META
template<typename... A> meta_any meta_type::invoke(id_type, meta_handle, A &&...) const ENTT_LITE_NO_META
inline id_type meta_type::id() const ENTT_LITE_NO_META inline const type_info &meta_type::info() const ENTT_LITE_NO_META
inline meta_range<meta_data> meta_type::data() const ENTT_LITE_NO_META inline meta_data meta_type::data(id_type) const ENTT_LITE_NO_META
inline bool meta_type::is_sequence_container() const ENTT_LITE_NO_META inline bool meta_type::is_associative_container() const ENTT_LITE_NO_META
inline bool meta_type::is_class() const ENTT_LITE_NO_META inline bool meta_type::is_arithmetic() const ENTT_LITE_NO_META
inline meta_type::operator bool() const ENTT_LITE_NO_META inline bool meta_type::operator==(const meta_type &) const ENTT_LITE_NO_META inline bool meta_type::operator!=(const meta_type &) const ENTT_LITE_NO_META
inline meta_any meta_type::construct() const ENTT_LITE_NO_META inline meta_any meta_type::from_void(void *) const ENTT_LITE_NO_META inline meta_any meta_type::from_void(const void *) const ENTT_LITE_NO_META
inline meta_handle::meta_handle(meta_any &) ENTT_LITE_NO_META inline meta_handle::meta_handle(const meta_any &) ENTT_LITE_NO_META inline meta_handle::operator bool() const ENTT_LITE_NO_META inline meta_any *meta_handle::operator->() ENTT_LITE_NO_META
inline meta_any meta_sequence_container::iterator::operator*() const ENTT_LITE_NO_META inline meta_sequence_container::iterator &meta_sequence_container::iterator::operator++() ENTT_LITE_NO_META
inline bool meta_sequence_container::iterator::operator!=(const iterator &) const ENTT_LITE_NO_META inline bool meta_sequence_container::iterator::operator==(const iterator &) const ENTT_LITE_NO_META inline meta_sequence_container::iterator::operator bool() const ENTT_LITE_NO_META
inline meta_type meta_sequence_container::value_type() const ENTT_LITE_NO_META inline std::size_t meta_sequence_container::size() const ENTT_LITE_NO_META inline bool meta_sequence_container::resize(std::size_t) ENTT_LITE_NO_META
inline bool meta_sequence_container::clear() ENTT_LITE_NO_META inline bool meta_sequence_container::reserve(std::size_t) ENTT_LITE_NO_META
inline meta_sequence_container::iterator meta_sequence_container::begin() ENTT_LITE_NO_META inline meta_sequence_container::iterator meta_sequence_container::end() ENTT_LITE_NO_META
inline meta_sequence_container::iterator meta_sequence_container::insert(iterator, meta_any) ENTT_LITE_NO_META inline meta_sequence_container::iterator meta_sequence_container::erase(iterator) ENTT_LITE_NO_META
inline meta_any meta_sequence_container::operator[](std::size_t) ENTT_LITE_NO_META inline meta_sequence_container::operator bool() const ENTT_LITE_NO_META
inline std::pair<meta_any, meta_any> meta_associative_container::iterator::operator*() const ENTT_LITE_NO_META inline meta_associative_container::iterator &meta_associative_container::iterator::operator++() ENTT_LITE_NO_META
inline bool meta_associative_container::iterator::operator!=(const iterator &) const ENTT_LITE_NO_META inline bool meta_associative_container::iterator::operator==(const iterator &) const ENTT_LITE_NO_META inline meta_associative_container::iterator::operator bool() const ENTT_LITE_NO_META
inline meta_type meta_associative_container::key_type() const ENTT_LITE_NO_META inline meta_type meta_associative_container::mapped_type() const ENTT_LITE_NO_META inline meta_type meta_associative_container::value_type() const ENTT_LITE_NO_META
inline std::size_t meta_associative_container::size() const ENTT_LITE_NO_META inline bool meta_associative_container::clear() ENTT_LITE_NO_META
inline meta_associative_container::iterator meta_associative_container::begin() ENTT_LITE_NO_META inline meta_associative_container::iterator meta_associative_container::end() ENTT_LITE_NO_META
inline bool meta_associative_container::insert(meta_any, meta_any) ENTT_LITE_NO_META inline std::size_t meta_associative_container::erase(meta_any) ENTT_LITE_NO_META inline meta_associative_container::iterator meta_associative_container::find(meta_any) ENTT_LITE_NO_META
inline meta_associative_container::operator bool() const ENTT_LITE_NO_META
inline meta_type meta_any::type() const ENTT_LITE_NO_META
inline meta_sequence_container meta_any::as_sequence_container() ENTT_LITE_NO_META inline meta_sequence_container meta_any::as_sequence_container() const ENTT_LITE_NO_META
inline meta_associative_container meta_any::as_associative_container() ENTT_LITE_NO_META inline meta_associative_container meta_any::as_associative_container() const ENTT_LITE_NO_META
inline meta_any meta_any::as_ref() ENTT_LITE_NO_META inline meta_any meta_any::as_ref() const ENTT_LITE_NO_META inline meta_any meta_any::get(id_type) const ENTT_LITE_NO_META
inline void *meta_any::data() ENTT_LITE_NO_META inline const void *meta_any::data() const ENTT_LITE_NO_META
inline meta_any::operator bool() const ENTT_LITE_NO_META inline bool meta_any::operator==(const meta_any &) const ENTT_LITE_NO_META inline bool meta_any::operator!=(const meta_any &) const ENTT_LITE_NO_META inline void meta_any::reset() ENTT_LITE_NO_META
inline meta_type meta_data::type() const ENTT_LITE_NO_META inline meta_any meta_data::get(meta_handle) const ENTT_LITE_NO_META
inline bool meta_data::is_const() const ENTT_LITE_NO_META inline bool meta_data::is_static() const ENTT_LITE_NO_META inline meta_data::operator bool() const ENTT_LITE_NO_META
template<typename T> meta_type resolve() noexcept ENTT_LITE_NO_META
inline meta_type resolve(id_type) noexcept ENTT_LITE_NO_META
inline meta_type resolve(const type_info &) noexcept ENTT_LITE_NO_META
struct as_is_t {}; struct as_ref_t {}; struct as_cref_t {}; struct as_void_t {};
// registration is accepted and ignored: edyn::attach registers entity-holding members for the (unreached) replication code
template<typename T> struct meta_factory {
    meta_factory type(id_type) { return *this; }
    template<auto D, typename Policy = as_is_t> meta_factory data(id_type) { return *this; }
    template<auto S, auto G, typename Policy = as_is_t> meta_factory data(id_type) { return *this; }
    template<auto F, typename Policy = as_is_t> meta_factory func(id_type) { return *this; }
    template<typename... A> meta_factory ctor() { return *this; }
    template<auto F, typename Policy = as_is_t> meta_factory ctor() { return *this; }
    template<typename B> meta_factory base() { return *this; }
    template<auto C> meta_factory conv() { return *this; }
    template<typename To> meta_factory conv() { return *this; }
};
template<typename T> meta_factory<T> meta() noexcept { return {}; }
template<typename T> void meta_reset() noexcept {}
inline void meta_reset() noexcept {}
template<typename T, typename... A> meta_any forward_as_meta(T &&) { ::entt::internal::no_meta(); }

} // namespace entt
