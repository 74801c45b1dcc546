// b2g_device_store.hpp — device shadows of block2 operator blocks (device-resident environments).
//
// The reference keeps renormalised environments in DataFrame stack 1, ONE partition at a time
// (core/allocator.hpp:617-660; MovingEnvironment::move_to loads the partition it needs and saves the new
// one, dmrg/moving_environment.hpp:1541-1575), so a host address inside that stack is reused by different
// partitions.  Device copies are therefore keyed by the OWNER OBJECT (the SparseMatrix of the operator,
// which lives as long as envs[i]->left / right does), never by host address: a shadow is valid while its
// owner is alive and still has the storage it had when the shadow was made.  The host binding collects the
// shadows of the operator tensors a call takes as arguments and hands the library the (host range ->
// device address) table for exactly that call (b2g_resident_map).
//
//   rotated environments (left_rotate / right_rotate results): written through - the device result is
//       also copied to the host block, which stays authoritative for the reference's own code
//       (partition files, intermediates, numerical_transform);
//   blocked operators (left_contract / right_contract results): device only - their host "storage" is
//       reserved address space without access rights, so that reference code reading them by accident
//       faults instead of computing with garbage; materialize() gives them real memory and content.
//
// Compiled together with block2's headers; talks to the CUDA library only through include/b2g.h.
#pragma once
#include "b2g.h"
#include "block2_core.hpp"
#include <chrono>
#include <stdexcept>
#include <sys/mman.h>
#include <unordered_map>

namespace b2g_host {

using namespace block2;

// One host address range for the blocks of one produced operator tensor.  accessible = false: reserved
// address space only (PROT_NONE, no memory behind it until materialised).
struct HostArena {
    double *base = nullptr;
    size_t bytes = 0;
    bool accessible = false;
    HostArena(size_t doubles, bool accessible) : accessible(accessible) {
        bytes = std::max<size_t>(doubles, 1) * sizeof(double);
        void *p = mmap(nullptr, bytes, accessible ? (PROT_READ | PROT_WRITE) : PROT_NONE,
                       MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
        if (p == MAP_FAILED)
            throw std::bad_alloc();
        base = (double *)p;
    }
    void make_accessible() {
        if (!accessible && mprotect(base, bytes, PROT_READ | PROT_WRITE) != 0)
            throw std::runtime_error("b2g: mprotect failed");
        accessible = true;
    }
    ~HostArena() { munmap(base, bytes); }
    HostArena(const HostArena &) = delete;
};

// Allocator handed to the SparseMatrix objects that live in a HostArena: the arena goes away when the
// last operator has been deallocated (SparseMatrix::deallocate drops its alloc pointer).
struct ArenaAllocator : Allocator<double> {
    shared_ptr<HostArena> arena;
    explicit ArenaAllocator(const shared_ptr<HostArena> &arena) : arena(arena) {}
    double *allocate(size_t) override { throw std::runtime_error("b2g: ArenaAllocator hands out no memory"); }
    void deallocate(void *, size_t) override {}
    double *reallocate(double *, size_t, size_t) override {
        throw std::runtime_error("b2g: ArenaAllocator cannot reallocate");
    }
};

struct DevBlock { // one device allocation holding the shadows made by one call
    b2g_context *ctx;
    double *base = nullptr;
    size_t doubles = 0;
    uint64_t last_use = 0;
    std::vector<const void *> keys; // shadows inside
    DevBlock(b2g_context *ctx, size_t n) : ctx(ctx), doubles(n) {
        if (b2g_malloc(ctx, std::max<size_t>(n, 2) * sizeof(double), (void **)&base) != 0)
            throw std::runtime_error(std::string("b2g_malloc: ") + b2g_last_error());
    }
    ~DevBlock() { b2g_free(ctx, base); }
    DevBlock(const DevBlock &) = delete;
};

struct Shadow {
    std::weak_ptr<void> owner;  // the SparseMatrix object
    const double *host;         // its data pointer and size when the shadow was made
    size_t doubles;
    const double *const *slot;  // &owner->data, &owner->total_memory (read only while the owner is alive)
    const size_t *size_slot;
    double *dev;
    shared_ptr<DevBlock> block;
    shared_ptr<HostArena> arena; // device-only blocks: the reserved host range (for materialize)
    bool host_valid;             // the host block holds the same values
    bool alive() const {
        std::shared_ptr<void> o = owner.lock();
        return o != nullptr && *slot == host && *size_slot == doubles;
    }
};

struct MapTable { // arguments of b2g_resident_map
    std::vector<const double *> host;
    std::vector<int64_t> doubles;
    std::vector<double *> dev;
    void add(const Shadow &s) { host.push_back(s.host), doubles.push_back((int64_t)s.doubles), dev.push_back(s.dev); }
};

struct DeviceStore {
    b2g_context *ctx;
    std::unordered_map<const void *, Shadow> shadows; // key = SparseMatrix*
    std::vector<shared_ptr<DevBlock>> blocks;
    size_t held = 0, peak = 0, budget = 0, evicted_bytes = 0, uploaded_bytes = 0, downloaded_bytes = 0;
    uint64_t clock = 1;
    explicit DeviceStore(b2g_context *ctx) : ctx(ctx) {
        int64_t f = 0, t = 0;
        b2g_mem_info(ctx, &f, &t);
        // what the shadows may hold: the rest is for the H.C plan (mirrored operands, W workspace, Davidson basis)
        const char *env = getenv("B2G_RESIDENT_GB");
        budget = env ? (size_t)(atof(env) * 1e9) : (size_t)(0.45 * (double)t);
    }
    void tick() { clock++; }
    void erase_key(const void *key) {
        auto it = shadows.find(key);
        if (it == shadows.end())
            return;
        shared_ptr<DevBlock> blk = it->second.block;
        shadows.erase(it);
        blk->keys.erase(std::remove(blk->keys.begin(), blk->keys.end(), key), blk->keys.end());
        if (blk->keys.empty())
            drop_block(blk);
    }
    void drop_block(const shared_ptr<DevBlock> &blk) {
        for (size_t i = 0; i < blocks.size(); i++)
            if (blocks[i] == blk) {
                held -= blk->doubles * sizeof(double);
                blocks.erase(blocks.begin() + (long)i);
                break;
            }
    }
    // forget the shadows whose owner died or gave its storage back
    void prune() {
        std::vector<const void *> dead;
        for (auto &kv : shadows)
            if (!kv.second.alive())
                dead.push_back(kv.first);
        for (const void *k : dead)
            erase_key(k);
        for (size_t i = blocks.size(); i-- > 0;) // blocks that never got a shadow and are not part of this call
            if (blocks[i]->keys.empty() && blocks[i]->last_use != clock) {
                held -= blocks[i]->doubles * sizeof(double);
                blocks.erase(blocks.begin() + (long)i);
            }
    }
    // least recently used blocks whose every shadow also lives on the host make room
    void make_room(size_t need_bytes) {
        while (held + need_bytes > budget) {
            shared_ptr<DevBlock> victim;
            for (auto &b : blocks) {
                if (b->last_use == clock)
                    continue; // part of the call in progress
                bool ok = true;
                for (const void *k : b->keys)
                    ok = ok && shadows.at(k).host_valid;
                if (ok && (victim == nullptr || b->last_use < victim->last_use))
                    victim = b;
            }
            if (victim == nullptr)
                break; // nothing evictable: let the allocation itself succeed or fail
            evicted_bytes += victim->doubles * sizeof(double);
            std::vector<const void *> keys = victim->keys;
            for (const void *k : keys)
                erase_key(k);
            drop_block(victim); // a block without shadows (nothing was added to it) is not dropped by erase_key
        }
    }
    // Memory pressure outside the store's own budget (the W workspace and the Davidson basis of a large site):
    // evict the least recently used written-through blocks until `fraction` of what is held is left, and hand
    // the unused pool memory back to the driver.  Returns the bytes released.
    // what is held, for diagnostics: (evictable bytes, pinned bytes = device-only blocks and blocks in use)
    std::pair<size_t, size_t> census() const {
        size_t ev = 0, pin = 0;
        for (auto &b : blocks) {
            bool ok = b->last_use != clock;
            for (const void *k : b->keys)
                ok = ok && shadows.at(k).host_valid;
            (ok ? ev : pin) += b->doubles * sizeof(double);
        }
        return std::make_pair(ev, pin);
    }
    size_t shrink(double fraction) {
        const size_t before = held, saved = budget;
        budget = (size_t)(fraction * (double)held);
        make_room(0);
        budget = saved;
        b2g_mem_trim(ctx);
        return before - held;
    }
    shared_ptr<DevBlock> new_block(size_t doubles, bool zero) {
        const bool prof = b2g_prof_enabled() != 0;
        auto t0 = std::chrono::steady_clock::now();
        auto lap = [&](const char *label) {
            if (prof) {
                const auto t1 = std::chrono::steady_clock::now();
                b2g_prof_record(label, std::chrono::duration<double>(t1 - t0).count());
                t0 = t1;
            }
        };
        make_room(doubles * sizeof(double));
        lap("store.new_block.make_room");
        shared_ptr<DevBlock> b;
        try {
            b = make_shared<DevBlock>(ctx, doubles);
        } catch (const std::runtime_error &) { // the device is full although the budget is not: make room, once
            if (shrink(0.5) == 0)
                throw;
            b = make_shared<DevBlock>(ctx, doubles);
        }
        lap("store.new_block.malloc");
        b->last_use = clock;
        if (zero && doubles != 0 && b2g_memset_zero(ctx, b->base, doubles * sizeof(double)) != 0)
            throw std::runtime_error(std::string("b2g_memset_zero: ") + b2g_last_error());
        lap("store.new_block.memset_issue");
        blocks.push_back(b);
        held += doubles * sizeof(double), peak = std::max(peak, held);
        return b;
    }
    template <typename SM>
    Shadow &add(const shared_ptr<DevBlock> &blk, const shared_ptr<SM> &m, double *dev, bool host_valid,
                const shared_ptr<HostArena> &arena = nullptr) {
        erase_key(m.get());
        Shadow s{std::weak_ptr<void>(std::shared_ptr<void>(m)), m->data, (size_t)m->total_memory,
                 (const double *const *)&m->data, (const size_t *)&m->total_memory, dev, blk, arena, host_valid};
        blk->keys.push_back(m.get());
        return shadows[m.get()] = s;
    }
    template <typename SM> Shadow *find(const shared_ptr<SM> &m) {
        if (m == nullptr)
            return nullptr;
        auto it = shadows.find(m.get());
        if (it == shadows.end())
            return nullptr;
        if (!it->second.alive() || it->second.owner.lock().get() != (void *)m.get()) {
            erase_key(m.get());
            return nullptr;
        }
        it->second.block->last_use = clock;
        return &it->second;
    }
    // shadows of every operator of a tensor (and, for a delayed tensor, of the two tensors behind it)
    template <typename S> void collect(const shared_ptr<OperatorTensor<S, double>> &opt, MapTable &tab) {
        if (opt == nullptr)
            return;
        for (auto &p : opt->ops)
            if (Shadow *s = find(p.second))
                tab.add(*s);
        if (opt->get_type() == OperatorTensorTypes::Delayed) {
            auto d = dynamic_pointer_cast<DelayedOperatorTensor<S, double>>(opt);
            collect<S>(d->lopt, tab), collect<S>(d->ropt, tab);
        }
    }
    // upload the operators of a tensor that have host content but no shadow yet (a partition that was
    // loaded from its file, intermediates the host computed): one block, read by the next two or three calls
    template <typename S> void ensure_shadows(const shared_ptr<OperatorTensor<S, double>> &opt, size_t min_doubles = 256) {
        if (opt == nullptr)
            return;
        std::vector<shared_ptr<SparseMatrix<S, double>>> todo;
        size_t total = 0;
        for (auto &p : opt->ops) {
            auto &m = p.second;
            if (m == nullptr || m->data == nullptr || m->total_memory < min_doubles || find(m) != nullptr)
                continue;
            bool dup = false;
            for (auto &q : todo)
                dup = dup || q == m;
            if (dup)
                continue;
            todo.push_back(m);
            total += (m->total_memory + 1) & ~(size_t)1;
        }
        if (todo.empty())
            return;
        shared_ptr<DevBlock> blk = new_block(total, false);
        std::vector<double *> dev;
        std::vector<const double *> host;
        std::vector<int64_t> n;
        size_t off = 0;
        for (auto &m : todo) {
            dev.push_back(blk->base + off), host.push_back(m->data), n.push_back((int64_t)m->total_memory);
            add(blk, m, blk->base + off, true);
            off += (m->total_memory + 1) & ~(size_t)1;
        }
        if (b2g_upload_blocks(ctx, (int64_t)dev.size(), dev.data(), host.data(), n.data()) != 0)
            throw std::runtime_error(std::string("b2g_upload_blocks: ") + b2g_last_error());
        uploaded_bytes += total * sizeof(double);
    }
    // give device-only operators of a tensor real host memory and their content (reference code is about
    // to read them: fallback paths, --verify)
    template <typename S> void materialize(const shared_ptr<OperatorTensor<S, double>> &opt) {
        if (opt == nullptr)
            return;
        std::vector<double *> host;
        std::vector<const double *> dev;
        std::vector<int64_t> n;
        for (auto &p : opt->ops)
            if (Shadow *s = find(p.second))
                if (!s->host_valid) {
                    if (s->arena != nullptr)
                        s->arena->make_accessible();
                    host.push_back((double *)s->host), dev.push_back(s->dev), n.push_back((int64_t)s->doubles);
                    s->host_valid = true;
                    downloaded_bytes += s->doubles * sizeof(double);
                }
        if (!host.empty() && b2g_download(ctx, (int64_t)host.size(), host.data(), dev.data(), n.data()) != 0)
            throw std::runtime_error(std::string("b2g_download: ") + b2g_last_error());
        if (opt->get_type() == OperatorTensorTypes::Delayed) {
            auto d = dynamic_pointer_cast<DelayedOperatorTensor<S, double>>(opt);
            materialize<S>(d->lopt), materialize<S>(d->ropt);
        }
    }
    void apply(const MapTable &tab) {
        // one operator may be reachable under several names: keep one entry per host block
        std::vector<size_t> idx(tab.host.size());
        for (size_t i = 0; i < idx.size(); i++)
            idx[i] = i;
        std::sort(idx.begin(), idx.end(), [&tab](size_t x, size_t y) { return tab.host[x] < tab.host[y]; });
        MapTable u;
        for (size_t z = 0; z < idx.size(); z++)
            if (z == 0 || tab.host[idx[z]] != tab.host[idx[z - 1]])
                u.host.push_back(tab.host[idx[z]]), u.doubles.push_back(tab.doubles[idx[z]]),
                    u.dev.push_back(tab.dev[idx[z]]);
        if (b2g_resident_map(ctx, (int64_t)u.host.size(), u.host.data(), u.doubles.data(), u.dev.data()) != 0)
            throw std::runtime_error(std::string("b2g_resident_map: ") + b2g_last_error());
    }
    void clear_map() { b2g_resident_map(ctx, 0, nullptr, nullptr, nullptr); }
    void drop_all() {
        clear_map();
        shadows.clear(), blocks.clear(), held = 0;
    }
};

} // namespace b2g_host
