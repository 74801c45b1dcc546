// b2g_adapter.hpp — the reference-side binding of libb2g.so.
//
// This header is compiled TOGETHER WITH block2's own headers (it is what a block2
// maintainer would add) and talks to the CUDA library only through
// the C ABI in include/b2g.h.  It installs the GPU executor behind the reference's own
// operator surface, without touching the MPO builder, the quantum-number bookkeeping or
// the sweep driver:
//
//   GPUTensorFunctions<S>  : TensorFunctions<S,double>   (core/tensor_functions.hpp:47)
//       operator()(b, c, scale)  -> b2g_seq_matvec     (was: opf->seq->operator()(b, c, scale), :59-62)
//       tensor_product_multiply  -> reference recording, then marks the device plan stale
//                                   (this is the call EffectiveHamiltonian::precompute() makes,
//                                    dmrg/effective_hamiltonian.hpp:226-246)
//       left_contract / right_contract (tensor_functions.hpp:2842-2885, 2941-2984)
//                                -> reference recording of the tensor_product list (SeqTypes::Auto),
//                                   executed by b2g_batch_execute instead of seq->auto_perform()
//       left_rotate / right_rotate (:2365-2403) -> recorded tensor_rotate list, b2g_pairs_execute
//       copy()                   -> keeps the dynamic type (EffectiveHamiltonian stores ptf->copy(), :137)
//   GPUDMRG<S>             : DMRG<S,double,double>        (dmrg/sweep_algorithm.hpp:71)
//       two_dot_eigs_and_perturb (virtual, :1183) -> H_eff built by the reference,
//                                   Davidson run device-resident by b2g_davidson
//
// The recording itself (OperatorFunctions::tensor_product_multiply /
// three_tensor_product_multiply -> BatchGEMMSeq::rotate / three_rotate) is the
// reference's, unchanged, so sector indexing and batch bookkeeping are bit-exact by
// construction; only the executor of the recorded list changes.
#pragma once
#include "b2g.h"
#include "block2_core.hpp"
#include "block2_dmrg.hpp"
#include "b2g_blocking_record.hpp"
#include "b2g_device_store.hpp"
#include "b2g_shm_comm.hpp"
#include <atomic>
#include <chrono>
#include <stdexcept>

namespace b2g_host {

using namespace block2;

// Density-matrix split: the reference diagonalises every block of the density matrix with LAPACK
// dsyev, one block per OpenMP thread (MovingEnvironment::truncate_density_matrix, dmrg/moving_environment.hpp:
// 3716-3790).  The build maps the reference's dsyev_ symbol to b2g_host_dsyev_ (blas_rename.h); with the hook armed
// (Session::gpu_split, b2g_dmrg --gpu-split) blocks of at least `min_n` rows go to b2g_syevd (cuSOLVER on the
// device, library-backed), everything else - workspace queries, small blocks, any failure - to the CPU routine.
struct SplitHook {
    static b2g_context *&ctx() {
        static b2g_context *c = nullptr;
        return c;
    }
    static int &min_n() {
        static int n = 256;
        return n;
    }
    static std::atomic<size_t> &calls_gpu() {
        static std::atomic<size_t> c{0};
        return c;
    }
    static std::atomic<size_t> &calls_cpu() {
        static std::atomic<size_t> c{0};
        return c;
    }
};

// B2G_PROF: wall-clock sections of the binding, accumulated by the library's profile (b2g_prof_record)
struct ProfLap {
    bool on;
    std::chrono::steady_clock::time_point t;
    ProfLap() : on(b2g_prof_enabled() != 0), t(std::chrono::steady_clock::now()) {}
    void lap(const char *label) {
        if (!on)
            return;
        const auto now = std::chrono::steady_clock::now();
        b2g_prof_record(label, std::chrono::duration<double>(now - t).count());
        t = now;
    }
};

struct Session {
    b2g_context *ctx = nullptr;
    std::atomic<bool> recording{false};
    double t_plan = 0, t_matvec = 0, t_precompute = 0, t_davidson = 0;
    double t_contract_alloc = 0, t_contract_ensure = 0, t_contract_exec = 0, t_rotate_exec = 0, t_rotate_alloc = 0;
    size_t n_plan = 0, n_matvec = 0, n_oom_retries = 0;
    // --verify: worst relative deviation ||sigma_gpu - sigma_cpu|| / ||sigma_cpu|| over all sites,
    // sigma_cpu from the reference's own BatchGEMMSeq::operator() on the same recorded list
    bool verify = false;
    double max_matvec_err = 0;
    size_t n_verified = 0;
    // renormalisation (left_rotate / right_rotate) on the device
    bool gpu_rotate = true;
    double t_rotate = 0, max_rotate_err = 0, t_rotate_download = 0;
    size_t n_rotate = 0, rotate_pairs = 0;
    double rotate_flops = 0;
    // blocking (left_contract / right_contract) on the device
    bool gpu_contract = true;
    double t_contract = 0, max_contract_err = 0, contract_kernel_ms = 0, contract_bytes = 0;
    double t_contract_record = 0, t_contract_plan = 0, t_contract_upload = 0, t_contract_download = 0;
    size_t n_contract = 0, contract_entries = 0;
    // intermediates / numerical_transform (lists of block additions) on the device
    bool gpu_iadd = true;
    double t_iadd = 0, max_iadd_err = 0;
    size_t n_iadd = 0, iadd_entries = 0;
    // dense eigenproblems of the density-matrix split on the device (cuSOLVER behind b2g_syevd); opt-in
    bool gpu_split = false;
    void arm_split(bool on) {
        gpu_split = on;
        SplitHook::ctx() = on ? ctx : nullptr;
    }
    // H_eff diagonal (tensor_product_diagonal) on the device
    bool gpu_diag = true;
    double t_diag = 0, max_diag_err = 0;
    size_t n_diag = 0, diag_entries = 0;
    // Device-resident environments (b2g_device_store.hpp).  host_mirror = true keeps every blocked operator
    // on the host as well (needed when reference code reads them: --verify, perturbative noise, ...).
    shared_ptr<DeviceStore> store;
    bool host_mirror = false;
    void *pinned = nullptr; // the DataFrame stacks, page-locked for direct DMA (pin_stacks)
    explicit Session(int device = 0) {
        if (b2g_context_create(device, &ctx) != 0)
            throw std::runtime_error(std::string("b2g_context_create: ") + b2g_last_error());
        store = make_shared<DeviceStore>(ctx);
    }
    // Both frame stacks come from one allocation (core/allocator.hpp:536): page-locking it lets the
    // write-through copies of the renormalised environments go by DMA straight into stack 1.
    void pin_stacks() {
        if (pinned != nullptr || frame_<double>() == nullptr || frame_<double>()->dallocs.size() < 2)
            return;
        // stack 1 only (the renormalised environments): page-locking touches every page, and most of the
        // main stack is never used when use_main_stack is false
        void *base = frame_<double>()->dallocs[1]->data;
        if (b2g_host_register(ctx, base, frame_<double>()->dallocs[1]->size * sizeof(double)) == 0)
            pinned = base;
    }
    ~Session() {
        if (SplitHook::ctx() == ctx)
            SplitHook::ctx() = nullptr;
        store->drop_all();
        store = nullptr;
        if (pinned != nullptr)
            b2g_host_unregister(ctx, pinned);
        b2g_context_destroy(ctx);
    }
    Session(const Session &) = delete;
};

// View of a BatchGEMM<double> as the C-ABI batch descriptor (same arrays, no copies).
inline b2g_batch as_b2g_batch(const BatchGEMM<double> &b) {
    static_assert(sizeof(CBLAS_TRANSPOSE) == sizeof(int32_t), "CBLAS_TRANSPOSE must be int-sized");
    static_assert(sizeof(MKL_INT) == sizeof(int32_t), "LP64 MKL_INT expected");
    for (size_t i = 0; i < b.gp.size(); i++)
        if (b.gp[i] != 1)
            throw std::runtime_error("b2g: grouped entries (gp != 1) are not part of the H.C replay list");
    if (b.acidxs.size() != 0)
        throw std::runtime_error("b2g: acidxs-tagged lists (partial expectation / complex) are not supported");
    b2g_batch r;
    r.count = (int64_t)b.gp.size();
    r.ta = (const int32_t *)b.ta.data(), r.tb = (const int32_t *)b.tb.data();
    r.m = b.m.data(), r.n = b.n.data(), r.k = b.k.data();
    r.lda = b.lda.data(), r.ldb = b.ldb.data(), r.ldc = b.ldc.data();
    r.alpha = b.alpha.data(), r.beta = b.beta.data();
    r.a = b.a.data(), r.b = b.b.data(), r.c = b.c.data();
    return r;
}

// Base = TensorFunctions<S,double> (serial) or ParallelTensorFunctions<S,double> (one process per
// GPU under ParallelRuleQC: the base keeps the reference's distributed blocking logic, the matvec
// and its sigma all-reduce run on the GPUs).
template <typename S, typename Base = TensorFunctions<S, double>> struct GPUTensorFunctions : Base {
    typedef double FL;
    using Base::opf;
    shared_ptr<Session> session;
    shared_ptr<ParallelRule<S, FL>> prule;
    mutable b2g_plan *plan = nullptr;
    mutable bool stale = true;
    mutable size_t csize = 0, vsize = 0;
    GPUTensorFunctions(const shared_ptr<OperatorFunctions<S, FL>> &opf, const shared_ptr<Session> &session)
        : Base(opf), session(session) {}
    GPUTensorFunctions(const shared_ptr<OperatorFunctions<S, FL>> &opf, const shared_ptr<ParallelRule<S, FL>> &rule,
                       const shared_ptr<Session> &session)
        : Base(opf, rule), session(session), prule(rule) {}
    ~GPUTensorFunctions() override { drop(); }
    void drop() const {
        if (plan != nullptr)
            b2g_plan_destroy(plan);
        plan = nullptr;
    }
    void forget() const { // after post_precompute(): the recorded H_eff is gone
        drop();
        plan_lopt = plan_ropt = nullptr;
    }
    shared_ptr<TensorFunctions<S, FL>> copy() const override { return make_copy((Base *)nullptr); }
    shared_ptr<TensorFunctions<S, FL>> make_copy(TensorFunctions<S, FL> *) const {
        return make_shared<GPUTensorFunctions<S, Base>>(opf->copy(), session);
    }
    shared_ptr<TensorFunctions<S, FL>> make_copy(ParallelTensorFunctions<S, FL> *) const {
        return make_shared<GPUTensorFunctions<S, Base>>(opf->copy(), prule, session);
    }
    // Top-level recording call of precompute(): run the reference's recorder, then invalidate
    // the device plan.  Nested calls (the per-term calls parallel_reduce makes on copies)
    // see `recording` set and only record.
    void tensor_product_multiply(const shared_ptr<OpExpr<S>> &expr, const shared_ptr<OpExpr<S>> &xexpr,
                                 const shared_ptr<OperatorTensor<S, FL>> &lopt,
                                 const shared_ptr<OperatorTensor<S, FL>> &ropt,
                                 const shared_ptr<SparseMatrix<S, FL>> &cmat,
                                 const shared_ptr<SparseMatrix<S, FL>> &vmat, S opdq,
                                 bool all_reduce) const override {
        const bool top = !session->recording.exchange(true);
        Base::tensor_product_multiply(expr, xexpr, lopt, ropt, cmat, vmat, opdq, all_reduce);
        if (top) {
            session->recording = false;
            if (cmat->data == nullptr && (opf->seq->mode & SeqTypes::Tasked)) {
                drop();
                stale = true;
                csize = cmat->total_memory, vsize = vmat->total_memory;
                plan_lopt = lopt, plan_ropt = ropt;
            }
        }
    }
    typedef unordered_map<shared_ptr<OpExpr<S>>, shared_ptr<SparseMatrix<S, FL>>> OpMap;
    DeviceStore &store() const { return *session->store; }
    static double rel_diff(const vector<vector<double>> &x, const vector<shared_ptr<SparseMatrix<S, FL>>> &ref) {
        double num = 0, den = 0;
        for (size_t z = 0; z < ref.size(); z++)
            for (size_t j = 0; j < ref[z]->total_memory; j++)
                num += (x[z][j] - ref[z]->data[j]) * (x[z][j] - ref[z]->data[j]), den += ref[z]->data[j] * ref[z]->data[j];
        return den > 0 ? sqrt(num / den) : sqrt(num);
    }
    // Renormalisation c = bra^T . a . ket of every operator of a block (core/tensor_functions.hpp:
    // 2365-2403).  The reference's own OperatorFunctions::tensor_rotate enumerates the sector blocks
    // and records one rotate() pair per block (operator_functions.hpp:175-210); in Auto mode nothing
    // is executed at record time, so the list is handed to b2g_pairs_execute instead of
    // seq->auto_perform().  The blocked operators are read from their device shadows, the results are
    // written into a fresh device block that shadows the new environment, and copied through to the
    // host blocks (DataFrame stack 1), which the reference saves as the partition file.
    void rotate_on_device(const shared_ptr<OperatorTensor<S, FL>> &a, const shared_ptr<SparseMatrix<S, FL>> &mpst_bra,
                          const shared_ptr<SparseMatrix<S, FL>> &mpst_ket, shared_ptr<OperatorTensor<S, FL>> &c,
                          const shared_ptr<Symbolic<S>> &names, bool trans) const {
        Timer t, td;
        t.get_time();
        ProfLap pl;
        store().tick(), store().prune();
        pl.lap("host.rotate.prune");
        auto &seq = opf->seq;
        if (seq->batch[0]->gp.size() != 0 || seq->batch[1]->gp.size() != 0)
            throw std::runtime_error("b2g: recorder not empty at rotate");
        // operators that are rotated: their host blocks are overwritten by the write-through copy, so
        // they are taken from the stack without the zero fill of SparseMatrix::allocate
        vector<shared_ptr<SparseMatrix<S, FL>>> outs;
        vector<shared_ptr<OpExpr<S>>> out_names;
        std::unordered_map<const void *, char> rotated;
        for (size_t i = 0; i < names->data.size(); i++)
            if (names->data[i]->get_type() != OpTypes::Zero) {
                auto pa = abs_value(names->data[i]);
                if (rotated.emplace(c->ops.at(pa).get(), 1).second)
                    outs.push_back(c->ops.at(pa)), out_names.push_back(pa);
            }
        for (auto &p : c->ops) { // allocation order of the stock method (stack allocator)
            auto &m = p.second;
            const size_t n = m->info->template get_total_memory<FL>();
            if (rotated.count(m.get()) && n != 0) {
                if (m->alloc == nullptr)
                    m->alloc = dalloc_<FL>();
                m->allocate(m->info, m->alloc->allocate(n));
            } else
                m->allocate(m->info);
        }
        pl.lap("host.rotate.host_alloc");
        size_t total = 0;
        for (auto &m : outs)
            total += (m->total_memory + 1) & ~(size_t)1;
        shared_ptr<DevBlock> blk = store().new_block(total, true);
        pl.lap("host.rotate.new_block");
        size_t off = 0;
        for (auto &m : outs) {
            if (m->total_memory != 0)
                store().add(blk, m, blk->base + off, false);
            off += (m->total_memory + 1) & ~(size_t)1;
        }
        pl.lap("host.rotate.add_shadows");
        {
            const double d = t.get_time();
            session->t_rotate_alloc += d, session->t_rotate += d;
        }
        const SeqTypes saved = seq->mode;
        seq->mode = SeqTypes::Auto; // record only
        for (size_t i = 0; i < out_names.size(); i++)
            opf->tensor_rotate(a->ops.at(out_names[i]), outs[i], mpst_bra, mpst_ket, trans);
        seq->mode = saved;
        pl.lap("host.rotate.record");
        if (session->verify || session->host_mirror)
            store().template materialize<S>(a);
        if (seq->batch[1]->gp.size() != 0) {
            b2g_batch b0 = as_b2g_batch(*seq->batch[0]), b1 = as_b2g_batch(*seq->batch[1]);
            MapTable tab;
            store().template collect<S>(a, tab);
            store().template collect<S>(c, tab);
            store().apply(tab);
            pl.lap("host.rotate.map");
            b2g_plan_stats st;
            Timer tx;
            tx.get_time();
            const int rc = b2g_pairs_execute(session->ctx, &b0, &b1, 0, &st);
            session->t_rotate_exec += tx.get_time();
            store().clear_map();
            pl.lap("host.rotate.pairs_execute");
            if (rc != 0)
                throw std::runtime_error(std::string("b2g_pairs_execute: ") + b2g_last_error());
            session->rotate_pairs += (size_t)st.pairs, session->rotate_flops += 2.0 * (double)st.nflop_mnk;
            seq->cumulative_nflop += (size_t)st.nflop_mnk;
        }
        // write through: the host blocks are what the reference saves, reloads and post-processes
        td.get_time();
        {
            vector<double *> host;
            vector<const double *> dev;
            vector<int64_t> n;
            for (auto &m : outs)
                if (Shadow *sh = store().find(m)) {
                    host.push_back(m->data), dev.push_back(sh->dev), n.push_back((int64_t)m->total_memory);
                    sh->host_valid = true;
                    store().downloaded_bytes += m->total_memory * sizeof(double);
                }
            if (!host.empty() && b2g_download(session->ctx, (int64_t)host.size(), host.data(), dev.data(), n.data()) != 0)
                throw std::runtime_error(std::string("b2g_download: ") + b2g_last_error());
        }
        session->t_rotate_download += td.get_time();
        pl.lap("host.rotate.write_through");
        if (session->verify && seq->batch[1]->gp.size() != 0) { // the reference executor on the same list
            vector<vector<double>> gpu;
            for (auto &m : outs) {
                gpu.emplace_back(m->data, m->data + m->total_memory);
                memset(m->data, 0, sizeof(double) * m->total_memory);
            }
            seq->mode = SeqTypes::Auto;
            seq->auto_perform();
            seq->mode = saved;
            session->max_rotate_err = max(session->max_rotate_err, rel_diff(gpu, outs));
            for (size_t z = 0; z < outs.size(); z++)
                memcpy(outs[z]->data, gpu[z].data(), sizeof(double) * outs[z]->total_memory);
        }
        seq->clear();
        pl.lap("host.rotate.clear");
        session->t_rotate += t.get_time(), session->n_rotate++;
    }
    int run_blocking_list(BatchGEMM<FL> &bt, b2g_blocking_stats &st) const {
        static_assert(sizeof(CBLAS_TRANSPOSE) == sizeof(int32_t) && sizeof(MKL_INT) == sizeof(int32_t), "");
        return b2g_batch_execute(session->ctx, (int64_t)bt.gp.size(), (const int32_t *)bt.ta.data(),
                                 (const int32_t *)bt.tb.data(), bt.m.data(), bt.n.data(), bt.k.data(), bt.alpha.data(),
                                 bt.a.data(), bt.lda.data(), bt.b.data(), bt.ldb.data(), bt.beta.data(), bt.c.data(),
                                 bt.ldc.data(), bt.gp.data(), B2G_OPERANDS_HOST, B2G_DST_ZERO, &st);
    }
    // Blocking c[op] = sum a[x] (x) b[y] (core/tensor_functions.hpp:2842-2885 / 2941-2984): the non-delayed,
    // not yet cached operators get a device block (and reserved host addresses, b2g_device_store.hpp),
    // every expression is walked in record-only mode and the terms run on the device, reading the
    // environment from its shadows and writing the blocked operators in place in HBM.
    void contract_on_device(const shared_ptr<OperatorTensor<S, FL>> &a, const shared_ptr<OperatorTensor<S, FL>> &b,
                            shared_ptr<OperatorTensor<S, FL>> &c, const shared_ptr<Symbolic<S>> &exprs,
                            const shared_ptr<Symbolic<S>> &names, OpNamesSet delayed, bool right) const {
        Timer t, tr;
        t.get_time();
        ProfLap pl;
        store().tick(), store().prune();
        pl.lap("host.contract.prune");
        auto &seq = opf->seq;
        if (seq->batch[0]->gp.size() != 0 || seq->batch[1]->gp.size() != 0)
            throw std::runtime_error("b2g: recorder not empty at contract");
        assert(exprs->data.size() == names->data.size());
        const SeqTypes saved = seq->mode;
        shared_ptr<GPUOperatorFunctions<S>> gopf = dynamic_pointer_cast<GPUOperatorFunctions<S>>(opf);
        if (gopf == nullptr)
            throw std::runtime_error("b2g: blocking on the device needs GPUOperatorFunctions (b2g_host::install)");
        const OpMap &lop = right ? b->ops : a->ops, &rop = right ? a->ops : b->ops;
        const bool mirror = session->verify || session->host_mirror;
        vector<size_t> todo;
        size_t total = 0;
        for (size_t i = 0; i < exprs->data.size(); i++) {
            shared_ptr<OpElement<S, FL>> cop = dynamic_pointer_cast<OpElement<S, FL>>(names->data[i]);
            shared_ptr<SparseMatrix<S, FL>> &m = c->ops.at(abs_value(names->data[i]));
            if (delayed(cop->name) || m->alloc != nullptr) // delayed, or the cached part
                continue;
            todo.push_back(i);
            total += (m->info->template get_total_memory<FL>() + 1) & ~(size_t)1;
        }
        if (mirror) // the cached part may have been produced device-only under the other policy
            store().template materialize<S>(c);
        Timer tx;
        tx.get_time();
        pl.lap("host.contract.todo");
        shared_ptr<HostArena> arena = make_shared<HostArena>(total, mirror);
        pl.lap("host.contract.arena");
        shared_ptr<DevBlock> blk = store().new_block(total, true);
        pl.lap("host.contract.new_block");
        vector<shared_ptr<SparseMatrix<S, FL>>> outs;
        size_t off = 0;
        for (size_t i : todo) {
            shared_ptr<SparseMatrix<S, FL>> m = c->ops.at(abs_value(names->data[i]));
            const size_t n = m->info->template get_total_memory<FL>();
            m->alloc = make_shared<ArenaAllocator>(arena);
            m->allocate(m->info, n == 0 ? nullptr : arena->base + off);
            if (n != 0)
                store().add(blk, m, blk->base + off, false, arena);
            outs.push_back(m);
            off += (n + 1) & ~(size_t)1;
        }
        session->t_contract_alloc += tx.get_time();
        pl.lap("host.contract.add_shadows");
        store().template ensure_shadows<S>(a); // the environment: a partition loaded from its file, intermediates
        session->t_contract_ensure += tx.get_time();
        pl.lap("host.contract.ensure_shadows");
        // record-only walk: the operators are walked by the operator-level threads, as the stock method
        // does with parallel_for; every thread has its own term vector, pre-sum recorder and temporaries,
        // and an operator is walked by one thread, so its terms stay in expression order
        vector<shared_ptr<SparseMatrix<S, FL>>> temps;
        vector<shared_ptr<OperatorFunctions<S, FL>>> pres; // pre-sum recorders (one per recording thread)
        tr.get_time();
        {
            gopf->collector->clear(), gopf->collector->active = true;
            const int nt = threading->activate_operator();
            if ((int)gopf->collector->per_thread.size() < nt)
                gopf->collector->per_thread.resize(nt);
            vector<shared_ptr<OperatorFunctions<S, FL>>> opfs(nt);
            vector<vector<shared_ptr<SparseMatrix<S, FL>>>> temps_t(nt);
            for (int k = 0; k < nt; k++) {
                opfs[k] = k == 0 ? opf : opf->copy();
                pres.push_back(make_shared<OperatorFunctions<S, FL>>(opf->cg));
                pres.back()->seq->mode = SeqTypes::Auto;
            }
            // SumProd temporaries (pre-sums without a stored intermediate) live on the device like the
            // blocked operators: reserved host addresses now, one device block once the walk is over
            const std::function<void(const shared_ptr<SparseMatrix<S, FL>> &, const shared_ptr<SparseMatrixInfo<S>> &)>
                alloc_tmp = [mirror](const shared_ptr<SparseMatrix<S, FL>> &m, const shared_ptr<SparseMatrixInfo<S>> &info) {
                    const size_t n = info->template get_total_memory<FL>();
                    shared_ptr<HostArena> ar = make_shared<HostArena>(n, mirror);
                    m->alloc = make_shared<ArenaAllocator>(ar);
                    m->allocate(info, n == 0 ? nullptr : ar->base);
                };
            std::exception_ptr err = nullptr;
#pragma omp parallel for schedule(dynamic) num_threads(nt)
            for (int z = 0; z < (int)todo.size(); z++) {
                const int tid = threading->get_thread_id();
                const size_t i = todo[z];
                try {
                    shared_ptr<OpElement<S, FL>> cop = dynamic_pointer_cast<OpElement<S, FL>>(names->data[i]);
                    record_blocking_expr<S>(opfs[tid], exprs->data[i] * ((FL)1.0 / cop->factor), lop, rop,
                                            c->ops.at(abs_value(names->data[i])), pres[tid], temps_t[tid], &alloc_tmp);
                } catch (...) {
#pragma omp critical
                    err = std::current_exception();
                }
            }
            threading->activate_normal();
            gopf->collector->active = false;
            if (err != nullptr)
                std::rethrow_exception(err);
            for (auto &tt : temps_t)
                temps.insert(temps.end(), tt.begin(), tt.end());
        }
        session->t_contract_record += tr.get_time();
        pl.lap("host.contract.record");
        auto account = [&](const b2g_blocking_stats &st) {
            session->contract_entries += (size_t)st.entries, session->contract_kernel_ms += st.kernel_ms;
            session->contract_bytes += (double)(st.bytes_in + st.bytes_out);
            session->t_contract_plan += st.plan_seconds, session->t_contract_upload += st.upload_seconds;
            session->t_contract_download += st.download_seconds;
            seq->cumulative_nflop += (size_t)st.nflop_mnk;
        };
        if (!temps.empty()) { // device block of the temporaries (zero: the pre-sum lists accumulate into it)
            size_t tt = 0;
            for (auto &m : temps)
                tt += (m->total_memory + 1) & ~(size_t)1;
            shared_ptr<DevBlock> tblk = store().new_block(tt, true);
            size_t toff = 0;
            for (auto &m : temps) {
                if (m->total_memory != 0)
                    store().add(tblk, m, tblk->base + toff, false,
                                dynamic_pointer_cast<ArenaAllocator>(m->alloc)->arena);
                toff += (m->total_memory + 1) & ~(size_t)1;
            }
        }
        MapTable tab;
        store().template collect<S>(a, tab);
        store().template collect<S>(c, tab);
        for (auto &m : temps)
            if (Shadow *sh = store().find(m))
                tab.add(*sh);
        store().apply(tab);
        pl.lap("host.contract.map");
        b2g_blocking_stats st;
        int rc = 0;
        bool chained = false;
        tx.get_time();
        for (auto &pre : pres) { // SumProd pre-sums into host temporaries (rare: stored intermediates cover most)
            if (pre->seq->batch[0]->gp.size() != 0)
                rc = 1, chained = true;
            else if (rc == 0 && pre->seq->batch[1]->gp.size() != 0) {
                rc = run_blocking_list(*pre->seq->batch[1], st);
                if (rc == 0)
                    account(st);
            }
        }
        pl.lap("host.contract.presums");
        vector<b2g_tp_term> &terms = gopf->collector->per_thread[0];
        for (size_t k = 1; k < gopf->collector->per_thread.size(); k++)
            terms.insert(terms.end(), gopf->collector->per_thread[k].begin(), gopf->collector->per_thread[k].end());
        pl.lap("host.contract.merge_terms");
        if (rc == 0 && terms.size() != 0) {
            rc = b2g_tensor_product_execute(session->ctx, (int64_t)terms.size(), terms.data(), B2G_OPERANDS_HOST,
                                            B2G_DST_ZERO, &st);
            if (rc == 0)
                account(st);
        }
        store().clear_map();
        session->t_contract_exec += tx.get_time();
        pl.lap("host.contract.tp_execute");
        gopf->collector->clear();
        seq->clear();
        for (auto &pre : pres)
            pre->seq->clear();
        pl.lap("host.contract.clear");
        if (rc != 0)
            throw std::runtime_error(chained ? std::string("b2g: blocking list has chained pairs")
                                             : std::string("b2g blocking: ") + b2g_last_error());
        if (mirror) {
            Timer tdn;
            tdn.get_time();
            store().template materialize<S>(c);
            session->t_contract_download += tdn.get_time();
        }
        if (session->verify && todo.size() != 0) {
            // the reference's own executor on the list its own recorder makes of the same expressions
            vector<vector<double>> gpu;
            for (auto &m : outs) {
                gpu.emplace_back(m->data, m->data + m->total_memory);
                if (m->total_memory != 0)
                    memset(m->data, 0, sizeof(double) * m->total_memory);
            }
            vector<shared_ptr<SparseMatrix<S, FL>>> temps2;
            shared_ptr<OperatorFunctions<S, FL>> pre2 = make_shared<OperatorFunctions<S, FL>>(opf->cg);
            shared_ptr<OperatorFunctions<S, FL>> stock = make_shared<OperatorFunctions<S, FL>>(opf->cg);
            stock->seq->mode = SeqTypes::Auto, pre2->seq->mode = SeqTypes::Auto;
            for (size_t i : todo) {
                shared_ptr<OpElement<S, FL>> cop = dynamic_pointer_cast<OpElement<S, FL>>(names->data[i]);
                record_blocking_expr<S>(stock, exprs->data[i] * ((FL)1.0 / cop->factor), lop, rop,
                                        c->ops.at(abs_value(names->data[i])), pre2, temps2);
            }
            pre2->seq->auto_perform();
            stock->seq->auto_perform();
            session->max_contract_err = max(session->max_contract_err, rel_diff(gpu, outs));
            for (size_t z = 0; z < outs.size(); z++)
                if (outs[z]->total_memory != 0)
                    memcpy(outs[z]->data, gpu[z].data(), sizeof(double) * outs[z]->total_memory);
            for (auto &m : temps2)
                m->deallocate();
        }
        for (auto &m : temps)
            m->deallocate();
        pl.lap("host.contract.tail");
        session->t_contract += t.get_time(), session->n_contract++;
    }
    bool on_device_ok() const { return !session->recording && !frame_<FL>()->use_main_stack; }
    void left_contract(const shared_ptr<OperatorTensor<S, FL>> &a, const shared_ptr<OperatorTensor<S, FL>> &b,
                       shared_ptr<OperatorTensor<S, FL>> &c, const shared_ptr<Symbolic<S>> &cexprs = nullptr,
                       OpNamesSet delayed = OpNamesSet()) const override {
        if (!session->gpu_contract || !on_device_ok() || a == nullptr || prule != nullptr)
            return Base::left_contract(a, b, c, cexprs, delayed);
        contract_on_device(a, b, c, cexprs == nullptr ? a->lmat * b->lmat : cexprs, c->lmat, delayed, false);
    }
    void right_contract(const shared_ptr<OperatorTensor<S, FL>> &a, const shared_ptr<OperatorTensor<S, FL>> &b,
                        shared_ptr<OperatorTensor<S, FL>> &c, const shared_ptr<Symbolic<S>> &cexprs = nullptr,
                        OpNamesSet delayed = OpNamesSet()) const override {
        if (!session->gpu_contract || !on_device_ok() || a == nullptr || prule != nullptr)
            return Base::right_contract(a, b, c, cexprs, delayed);
        contract_on_device(a, b, c, cexprs == nullptr ? b->rmat * a->rmat : cexprs, c->rmat, delayed, true);
    }
    void left_rotate(const shared_ptr<OperatorTensor<S, FL>> &a, const shared_ptr<SparseMatrix<S, FL>> &mpst_bra,
                     const shared_ptr<SparseMatrix<S, FL>> &mpst_ket,
                     shared_ptr<OperatorTensor<S, FL>> &c) const override {
        if (!session->gpu_rotate || session->recording || prule != nullptr) { // under a parallel rule: the reference's
            // distributed renormalisation (only the operators this rank holds, parallel_tensor_functions.hpp:881-907)
            store().template materialize<S>(a);
            return Base::left_rotate(a, mpst_bra, mpst_ket, c);
        }
        rotate_on_device(a, mpst_bra, mpst_ket, c, a->lmat, false);
    }
    void right_rotate(const shared_ptr<OperatorTensor<S, FL>> &a, const shared_ptr<SparseMatrix<S, FL>> &mpst_bra,
                      const shared_ptr<SparseMatrix<S, FL>> &mpst_ket,
                      shared_ptr<OperatorTensor<S, FL>> &c) const override {
        if (!session->gpu_rotate || session->recording || prule != nullptr) { // under a parallel rule: the reference's
            // distributed renormalisation (only the operators this rank holds, parallel_tensor_functions.hpp:881-907)
            store().template materialize<S>(a);
            return Base::right_rotate(a, mpst_bra, mpst_ket, c);
        }
        rotate_on_device(a, mpst_bra, mpst_ket, c, a->rmat, true);
    }
    // H_eff diagonal (core/tensor_functions.hpp:2027-2182): the stock walk runs unchanged, but the
    // GPUOperatorFunctions of every walking thread records its tensor_product_diagonal /
    // three_tensor_product_diagonal calls into a recorder of its own (b2g_blocking_record.hpp) instead of
    // the shared sequence the stock method would execute on the host; the recorded k = 1 outer products
    // diag[i, j] += f * A[i, i] * B[j, j] then run on the device, reading the operator diagonals from the
    // shadows, and only the diagonal itself (|psi| doubles) comes back.
    void tensor_product_diagonal(const shared_ptr<OpExpr<S>> &expr, const shared_ptr<OpExpr<S>> &xexpr,
                                 const shared_ptr<OperatorTensor<S, FL>> &lopt,
                                 const shared_ptr<OperatorTensor<S, FL>> &ropt,
                                 const shared_ptr<SparseMatrix<S, FL>> &mat, S opdq) const override {
        shared_ptr<GPUOperatorFunctions<S>> gopf = dynamic_pointer_cast<GPUOperatorFunctions<S>>(opf);
        const bool top = gopf != nullptr && session->gpu_diag && !gopf->collector->diag_active &&
                         !session->recording && (opf->seq->mode & SeqTypes::Tasked);
        if (!top) {
            if (gopf == nullptr || !gopf->collector->diag_active)
                store().template materialize<S>(lopt), store().template materialize<S>(ropt);
            return Base::tensor_product_diagonal(expr, xexpr, lopt, ropt, mat, opdq);
        }
        Timer t;
        t.get_time();
        store().tick(), store().prune();
        gopf->collector->begin_diag();
        try {
            Base::tensor_product_diagonal(expr, xexpr, lopt, ropt, mat, opdq); // records; executes nothing
        } catch (...) {
            gopf->collector->end_diag();
            throw;
        }
        gopf->collector->diag_active = false;
        MapTable tab;
        store().template collect<S>(lopt, tab);
        store().template collect<S>(ropt, tab);
        store().apply(tab);
        int rc = 0;
        b2g_blocking_stats st;
        for (auto &ds : gopf->collector->diag_seqs)
            if (rc == 0 && ds != nullptr && ds->batch[1]->gp.size() != 0) {
                rc = run_blocking_list(*ds->batch[1], st);
                session->diag_entries += (size_t)st.entries;
                opf->seq->cumulative_nflop += (size_t)st.nflop_mnk;
            }
        store().clear_map();
        if (rc != 0) {
            gopf->collector->end_diag();
            throw std::runtime_error(std::string("b2g diagonal: ") + b2g_last_error());
        }
        if (session->verify) { // the reference executor on the same recorded lists
            vector<double> gpu(mat->data, mat->data + mat->total_memory);
            memset(mat->data, 0, sizeof(double) * mat->total_memory);
            store().template materialize<S>(lopt), store().template materialize<S>(ropt);
            for (auto &ds : gopf->collector->diag_seqs)
                if (ds != nullptr && ds->batch[1]->gp.size() != 0)
                    ds->auto_perform();
            double num = 0, den = 0;
            for (size_t j = 0; j < mat->total_memory; j++)
                num += (gpu[j] - mat->data[j]) * (gpu[j] - mat->data[j]), den += mat->data[j] * mat->data[j];
            session->max_diag_err = max(session->max_diag_err, den > 0 ? sqrt(num / den) : sqrt(num));
            memcpy(mat->data, gpu.data(), sizeof(double) * mat->total_memory);
        }
        gopf->collector->end_diag();
        session->t_diag += t.get_time(), session->n_diag++;
    }
    // intermediates (pre-sums of complementary operators stored with the environment) and numerical_transform
    // (normal -> complementary operators at the middle site), core/tensor_functions.hpp:2404-2517: plain
    // lists of block additions c += f * op(b).  The stock walk runs unchanged (it also creates the new
    // operators), its OperatorFunctions::iadd calls are recorded per walking thread, and the lists run on
    // the device: sources are the shadows of the environment the rotation just produced, the results get
    // shadows of their own and are written through to the host blocks (DataFrame stack 1) the reference saves.
    template <typename Walk> void iadd_walk_on_device(const shared_ptr<OperatorTensor<S, FL>> &a, Walk walk) const {
        shared_ptr<GPUOperatorFunctions<S>> gopf = dynamic_pointer_cast<GPUOperatorFunctions<S>>(opf);
        Timer t;
        t.get_time();
        store().tick(), store().prune();
        gopf->collector->begin_iadd();
        try {
            walk(); // allocates the new operators (zero) and records; executes nothing
        } catch (...) {
            gopf->collector->end_iadd();
            throw;
        }
        gopf->collector->iadd_active = false;
        // operators the lists write: those of `a` that hold an output address
        vector<const double *> outp;
        for (auto &q : gopf->collector->iadd_seqs)
            if (q->batch[0]->gp.size() != 0) {
                gopf->collector->end_iadd();
                throw std::runtime_error("b2g: iadd list has chained pairs");
            } else
                outp.insert(outp.end(), q->batch[1]->c.begin(), q->batch[1]->c.end());
        vector<b2g_tp_term> terms;
        for (auto &tv : gopf->collector->iadd_terms)
            terms.insert(terms.end(), tv.begin(), tv.end());
        for (auto &tm : terms)
            outp.push_back(tm.c);
        std::sort(outp.begin(), outp.end());
        vector<shared_ptr<SparseMatrix<S, FL>>> outs;
        std::unordered_map<const void *, char> seen;
        size_t total = 0;
        for (auto &p : a->ops) {
            auto &m = p.second;
            if (m == nullptr || m->data == nullptr || m->total_memory == 0 || seen.count(m.get()))
                continue;
            auto it = std::lower_bound(outp.begin(), outp.end(), (const double *)m->data);
            if (it != outp.end() && *it < m->data + m->total_memory) {
                if (store().find(m) != nullptr) { // an output that already has a shadow: its host copy is current
                    gopf->collector->end_iadd();   // (written through), but the list would change it
                    throw std::runtime_error("b2g: iadd list writes an operator that is already resident");
                }
                seen.emplace(m.get(), 1), outs.push_back(m);
                total += (m->total_memory + 1) & ~(size_t)1;
            }
        }
        shared_ptr<DevBlock> blk = store().new_block(total, true);
        size_t off = 0;
        for (auto &m : outs) {
            store().add(blk, m, blk->base + off, false);
            off += (m->total_memory + 1) & ~(size_t)1;
        }
        MapTable tab;
        store().template collect<S>(a, tab);
        store().apply(tab);
        int rc = 0;
        b2g_blocking_stats st;
        for (auto &q : gopf->collector->iadd_seqs)
            if (rc == 0 && q->batch[1]->gp.size() != 0) {
                rc = run_blocking_list(*q->batch[1], st);
                session->iadd_entries += (size_t)st.entries;
                opf->seq->cumulative_nflop += (size_t)st.nflop_mnk;
            }
        if (rc == 0 && !terms.empty()) {
            rc = b2g_tensor_product_execute(session->ctx, (int64_t)terms.size(), terms.data(), B2G_OPERANDS_HOST,
                                            B2G_DST_ZERO, &st);
            session->iadd_entries += (size_t)st.entries;
            opf->seq->cumulative_nflop += (size_t)st.nflop_mnk;
        }
        store().clear_map();
        if (rc != 0) {
            gopf->collector->end_iadd();
            throw std::runtime_error(std::string("b2g iadd lists: ") + b2g_last_error());
        }
        vector<double *> host;
        vector<const double *> dev;
        vector<int64_t> n;
        for (auto &m : outs)
            if (Shadow *sh = store().find(m)) {
                host.push_back(m->data), dev.push_back(sh->dev), n.push_back((int64_t)m->total_memory);
                sh->host_valid = true;
                store().downloaded_bytes += m->total_memory * sizeof(double);
            }
        if (!host.empty() && b2g_download(session->ctx, (int64_t)host.size(), host.data(), dev.data(), n.data()) != 0) {
            gopf->collector->end_iadd();
            throw std::runtime_error(std::string("b2g_download: ") + b2g_last_error());
        }
        if (session->verify) { // the reference executor on the same recorded lists
            vector<vector<double>> gpu;
            for (auto &m : outs) {
                gpu.emplace_back(m->data, m->data + m->total_memory);
                memset(m->data, 0, sizeof(double) * m->total_memory);
            }
            for (auto &q : gopf->collector->iadd_seqs)
                if (q->batch[1]->gp.size() != 0)
                    q->auto_perform();
            for (auto &tm : terms) // the eager reference routine on every block descriptor
                GMatrixFunctions<FL>::iadd(GMatrix<FL>(tm.c, tm.conja ? tm.an : tm.am, tm.conja ? tm.am : tm.an),
                                           GMatrix<FL>((FL *)tm.a, tm.am, tm.an), tm.scale, tm.conja != 0);
            session->max_iadd_err = max(session->max_iadd_err, rel_diff(gpu, outs));
            for (size_t z = 0; z < outs.size(); z++)
                memcpy(outs[z]->data, gpu[z].data(), sizeof(double) * outs[z]->total_memory);
        }
        gopf->collector->end_iadd();
        session->t_iadd += t.get_time(), session->n_iadd++;
    }
    bool iadd_on_device_ok() const {
        shared_ptr<GPUOperatorFunctions<S>> gopf = dynamic_pointer_cast<GPUOperatorFunctions<S>>(opf);
        return gopf != nullptr && session->gpu_iadd && !session->recording && !gopf->collector->iadd_active &&
               prule == nullptr && (opf->seq->mode == SeqTypes::Tasked || opf->seq->mode == SeqTypes::None);
    }
    void intermediates(const shared_ptr<Symbolic<S>> &names, const shared_ptr<Symbolic<S>> &exprs,
                       const shared_ptr<OperatorTensor<S, FL>> &a, bool left) const override {
        if (!iadd_on_device_ok())
            return Base::intermediates(names, exprs, a, left);
        iadd_walk_on_device(a, [&]() { Base::intermediates(names, exprs, a, left); });
    }
    void numerical_transform(const shared_ptr<OperatorTensor<S, FL>> &a, const shared_ptr<Symbolic<S>> &names,
                             const shared_ptr<Symbolic<S>> &exprs) const override {
        if (!iadd_on_device_ok())
            return Base::numerical_transform(a, names, exprs);
        iadd_walk_on_device(a, [&]() { Base::numerical_transform(a, names, exprs); });
    }
    // Everything else that reads operator tensors on the host gets real host copies first.
    void tensor_product_partial_multiply(const shared_ptr<OpExpr<S>> &expr, const shared_ptr<OpExpr<S>> &xexpr,
                                         const shared_ptr<OperatorTensor<S, FL>> &lopt,
                                         const shared_ptr<OperatorTensor<S, FL>> &ropt, bool trace_right,
                                         const shared_ptr<SparseMatrix<S, FL>> &cmat,
                                         const vector<pair<uint8_t, S>> &psubsl,
                                         const vector<vector<shared_ptr<typename SparseMatrixInfo<S>::ConnectionInfo>>> &cinfos,
                                         const vector<S> &vdqs, const shared_ptr<SparseMatrixGroup<S, FL>> &vmats,
                                         int &vidx, int tvidx, bool do_reduce) const override {
        if (!session->recording)
            store().template materialize<S>(lopt), store().template materialize<S>(ropt);
        Base::tensor_product_partial_multiply(expr, xexpr, lopt, ropt, trace_right, cmat, psubsl, cinfos, vdqs, vmats, vidx,
                                              tvidx, do_reduce);
    }
    mutable shared_ptr<OperatorTensor<S, FL>> plan_lopt, plan_ropt; // operator tensors of the recorded H_eff
    void build_plan() const {
        Timer t;
        t.get_time();
        drop();
        auto &seq = opf->seq;
        ProfLap pl;
        b2g_batch b0 = as_b2g_batch(*seq->batch[0]), b1 = as_b2g_batch(*seq->batch[1]);
        store().tick(), store().prune();
        MapTable tab; // environments and blocked operators of this H_eff are read in place from HBM
        store().template collect<S>(plan_lopt, tab);
        store().template collect<S>(plan_ropt, tab);
        store().apply(tab);
        pl.lap("host.plan.map");
        const int rc = b2g_plan_create(session->ctx, &b0, &b1, (int64_t)seq->max_work, (int64_t)csize, (int64_t)vsize,
                                       B2G_OPERANDS_HOST, &plan);
        store().clear_map();
        if (rc != 0)
            throw std::runtime_error(std::string("b2g_plan_create: ") + b2g_last_error());
        stale = false;
        session->t_plan += t.get_time(), session->n_plan++;
    }
    b2g_plan *get_plan() const {
        if (stale || plan == nullptr)
            build_plan();
        return plan;
    }
    // sigma += scale * H.c  (host buffers, same contract as BatchGEMMSeq::operator())
    void operator()(const GMatrix<FL> &b, const GMatrix<FL> &c, FL scale = 1.0) override {
        if (!(opf->seq->mode & SeqTypes::Tasked))
            throw std::runtime_error("b2g: GPUTensorFunctions needs SeqTypes::Tasked (or SimpleTasked)");
        // a rank without terms at this site still joins the sigma all-reduce (the reference's
        // ParallelTensorFunctions::operator() always does, parallel_tensor_functions.hpp:51-55)
        if (opf->seq->batch[0]->gp.size() == 0 && prule == nullptr)
            return;
        Timer t;
        t.get_time();
        if (b2g_seq_matvec(get_plan(), b.data, c.data, scale) != 0)
            throw std::runtime_error(std::string("b2g_seq_matvec: ") + b2g_last_error());
        // keep the reference's FLOP accounting (batch_gemm.hpp:1687-1688)
        opf->seq->cumulative_nflop += opf->seq->batch[0]->nflop + opf->seq->batch[1]->nflop;
        session->t_matvec += t.get_time(), session->n_matvec++;
    }
};

// DMRG with the Davidson solver device-resident. Everything that is not the plain
// ground-state path falls through to the reference implementation.
template <typename S> struct GPUDMRG : DMRG<S, double, double> {
    typedef DMRG<S, double, double> Base;
    typedef typename Base::FPLS FPLS;
    using Base::me;
    bool device_davidson = true;
    GPUDMRG(const shared_ptr<MovingEnvironment<S, double, double>> &me, const vector<ubond_t> &bond_dims,
            const vector<double> &noises)
        : Base(me, bond_dims, noises) {}
    tuple<FPLS, int, size_t, double>
    two_dot_eigs_and_perturb(const bool forward, const int i, const double davidson_conv_thrd, const double noise,
                             shared_ptr<SparseMatrixGroup<S, double>> &pket) override {
        const bool plain = device_davidson && !this->state_specific && this->projection_weights.size() == 0 &&
                           this->metric_me == nullptr && this->context_ket == nullptr &&
                           this->davidson_type == DavidsonTypes::Normal && this->eff_kernel == nullptr &&
                           !((this->noise_type & NoiseTypes::Perturbative) && noise != 0);
        if (!plain) {
            // the stock solver / perturbative noise read operator blocks on the host: from here on every
            // blocked operator is kept on the host as well
            auto g0 = dynamic_pointer_cast<GPUTensorFunctions<S>>(me->mpo->tf);
            auto g1 = dynamic_pointer_cast<GPUTensorFunctions<S, ParallelTensorFunctions<S, double>>>(me->mpo->tf);
            if (g0 != nullptr)
                g0->session->host_mirror = true;
            if (g1 != nullptr)
                g1->session->host_mirror = true;
            return Base::two_dot_eigs_and_perturb(forward, i, davidson_conv_thrd, noise, pket);
        }
        Timer t;
        t.get_time();
        ProfLap pl0;
        shared_ptr<EffectiveHamiltonian<S, double>> h_eff =
            me->eff_ham(FuseTypes::FuseLR, forward, true, me->bra->tensors[i], me->ket->tensors[i]);
        pl0.lap("host.eigs.eff_ham");
        this->sweep_max_eff_ham_size = max(this->sweep_max_eff_ham_size, h_eff->op->get_total_memory());
        this->sweep_max_eff_wfn_size = max(this->sweep_max_eff_wfn_size, (size_t)h_eff->ket->total_memory);
        this->teff += t.get_time();
        auto gtf_s = dynamic_pointer_cast<GPUTensorFunctions<S>>(h_eff->tf);
        auto gtf_p = dynamic_pointer_cast<GPUTensorFunctions<S, ParallelTensorFunctions<S, double>>>(h_eff->tf);
        if (gtf_s == nullptr && gtf_p == nullptr)
            throw std::runtime_error("b2g: GPUDMRG needs mpo->tf to be a GPUTensorFunctions");
        if (gtf_p != nullptr)
            return eigs_on_device(gtf_p, h_eff, davidson_conv_thrd, t);
        return eigs_on_device(gtf_s, h_eff, davidson_conv_thrd, t);
    }
    template <typename GTF>
    tuple<FPLS, int, size_t, double> eigs_on_device(const shared_ptr<GTF> &gtf,
                                                    const shared_ptr<EffectiveHamiltonian<S, double>> &h_eff,
                                                    const double davidson_conv_thrd, Timer &t) {
        frame_<double>()->activate(0);
        Timer tq;
        tq.get_time();
        ProfLap pl;
        h_eff->precompute();
        gtf->session->t_precompute += tq.get_time();
        pl.lap("host.eigs.precompute");
        if (gtf->session->verify && h_eff->tf->opf->seq->batch[0]->gp.size() != 0) {
            const size_t n = h_eff->ket->total_memory;
            vector<double> x(n), y_gpu(n, 0.0), y_cpu(n, 0.0);
            Random::fill<double>(x.data(), n);
            GMatrix<double> xm(x.data(), (MKL_INT)n, 1);
            (*gtf)(xm, GMatrix<double>(y_gpu.data(), (MKL_INT)n, 1), 1.0);
            gtf->store().template materialize<S>(gtf->plan_lopt), gtf->store().template materialize<S>(gtf->plan_ropt);
            h_eff->tf->opf->seq->operator()(xm, GMatrix<double>(y_cpu.data(), (MKL_INT)n, 1), 1.0);
            if (me->para_rule != nullptr) // the GPU result is already summed over ranks (NCCL)
                me->para_rule->comm->allreduce_sum(y_cpu.data(), n);
            double num = 0, den = 0;
            for (size_t j = 0; j < n; j++)
                num += (y_gpu[j] - y_cpu[j]) * (y_gpu[j] - y_cpu[j]), den += y_cpu[j] * y_cpu[j];
            const double err = den > 0 ? sqrt(num / den) : sqrt(num);
            gtf->session->max_matvec_err = max(gtf->session->max_matvec_err, err);
            gtf->session->n_verified++;
        }
        double e = 0;
        int ndav = 0;
        size_t nflop = 0;
        // under a parallel rule a rank without terms at this site still runs the solver: every rank joins the
        // sigma all-reduces of the replicated Davidson iteration
        if (h_eff->tf->opf->seq->batch[0]->gp.size() != 0 || me->para_rule != nullptr) {
            pl.lap("host.eigs.verify");
            // The plan workspace (W panels, partial tiles) and the Davidson basis are allocated outside the budget
            // of the resident environments.  When the device is full, the least recently used written-through
            // environments go (they are re-read from their host blocks on demand) and the site is tried again.
            for (int attempt = 0;; attempt++) {
                try {
                    gtf->get_plan();
                    pl.lap("host.eigs.plan");
                    tq.get_time();
                    if (b2g_davidson(gtf->get_plan(), h_eff->diag->data, h_eff->ket->data, davidson_conv_thrd,
                                     this->davidson_rel_conv_thrd, this->davidson_max_iter, this->davidson_soft_max_iter,
                                     this->davidson_def_min_size, this->davidson_def_max_size, &e, &ndav) != 0)
                        throw std::runtime_error(std::string("b2g_davidson: ") + b2g_last_error());
                    break;
                } catch (const std::runtime_error &err) {
                    const bool oom = std::string(err.what()).find("out of memory") != std::string::npos ||
                                     std::string(err.what()).find("failed") != std::string::npos;
                    b2g_plan_stats pst;
                    memset(&pst, 0, sizeof(pst));
                    if (gtf->plan != nullptr)
                        b2g_plan_get_stats(gtf->plan, &pst);
                    gtf->drop(), gtf->stale = true;
                    gtf->store().tick(); // nothing of the failed attempt is "in use"
                    int64_t mf = 0, mt = 0;
                    b2g_mem_info(gtf->session->ctx, &mf, &mt);
                    const std::pair<size_t, size_t> cs = gtf->store().census();
                    fprintf(stderr,
                            "[b2g] eigs attempt %d failed: %s\n[b2g]   |psi| %zu doubles, plan: workspace %.2f GB, mirrored "
                            "operands %.2f GB; store: %.2f GB evictable + %.2f GB device-only of budget %.2f GB in %zu "
                            "blocks; device: %.2f GB free of %.2f GB\n",
                            attempt, err.what(), (size_t)h_eff->ket->total_memory, pst.workspace_doubles * 8e-9,
                            pst.mirrored_doubles * 8e-9, cs.first * 1e-9, cs.second * 1e-9, gtf->store().budget * 1e-9,
                            gtf->store().blocks.size(), mf * 1e-9, mt * 1e-9);
                    // first retry: half of the evictable environments go; second: all of them
                    if (!oom || attempt >= 2 || gtf->store().shrink(attempt == 0 ? 0.5 : 0.0) == 0)
                        throw;
                    gtf->session->n_oom_retries++;
                }
            }
            gtf->session->t_davidson += tq.get_time();
            pl.lap("host.eigs.davidson");
            if (b2g_prof_enabled()) { // memory of the site, beside the library's own per-site Davidson line
                b2g_plan_stats pst;
                memset(&pst, 0, sizeof(pst));
                b2g_plan_get_stats(gtf->get_plan(), &pst);
                int64_t mf = 0, mt = 0;
                b2g_mem_info(gtf->session->ctx, &mf, &mt);
                const std::pair<size_t, size_t> cs = gtf->store().census();
                fprintf(stderr, "[b2g] site memory: plan workspace %.2f GB, mirrored operands %.2f GB; store %.2f GB evictable "
                                "+ %.2f GB device-only / in use; device %.2f GB free (pool included) of %.2f GB\n",
                        pst.workspace_doubles * 8e-9, pst.mirrored_doubles * 8e-9, cs.first * 1e-9, cs.second * 1e-9,
                        mf * 1e-9, mt * 1e-9);
            }
            nflop = (size_t)ndav * (h_eff->tf->opf->seq->batch[0]->nflop + h_eff->tf->opf->seq->batch[1]->nflop);
        }
        h_eff->post_precompute();
        pl.lap("host.eigs.post_precompute");
        gtf->forget();
        pl.lap("host.eigs.forget");
        double tdav = t.get_time();
        this->teig += tdav;
        h_eff->deallocate();
        pl.lap("host.eigs.heff_deallocate");
        return make_tuple((FPLS)e, ndav, nflop, tdav);
    }
};

// Install the executor on an MPO (after simplification): mpo->tf is the only seam
// MovingEnvironment and EffectiveHamiltonian use (moving_environment.hpp:329,362,2173).
template <typename S>
inline shared_ptr<Session> install(const shared_ptr<MPO<S, double>> &mpo, int device = 0) {
    shared_ptr<Session> session = make_shared<Session>(device);
    shared_ptr<GPUOperatorFunctions<S>> gopf =
        make_shared<GPUOperatorFunctions<S>>(mpo->tf->opf->cg, make_shared<TermCollector>());
    gopf->seq = mpo->tf->opf->seq;
    mpo->tf = make_shared<GPUTensorFunctions<S>>(gopf, session);
    return session;
}

// Multi-GPU: mpo is a ParallelMPO over ParallelRuleQC (one process per GPU).  The NCCL communicator
// of the session is created from an id that rank 0 broadcasts through the host communicator.
template <typename S>
inline shared_ptr<Session> install_parallel(const shared_ptr<MPO<S, double>> &mpo, int device = 0) {
    shared_ptr<ParallelMPO<S, double>> pmpo = dynamic_pointer_cast<ParallelMPO<S, double>>(mpo);
    shared_ptr<ClassicParallelMPO<S, double>> cmpo = dynamic_pointer_cast<ClassicParallelMPO<S, double>>(mpo);
    if (pmpo == nullptr && cmpo == nullptr)
        throw std::runtime_error("b2g: install_parallel needs a ParallelMPO or a ClassicParallelMPO");
    shared_ptr<ParallelRule<S, double>> prule = pmpo != nullptr ? pmpo->rule : cmpo->rule;
    shared_ptr<Session> session = make_shared<Session>(device);
    auto comm = prule->comm;
    char id[128];
    if (comm->rank == comm->root && b2g_comm_unique_id(id) != 0)
        throw std::runtime_error(std::string("b2g_comm_unique_id: ") + b2g_last_error());
    static_assert(sizeof(long long int) == 8, "");
    comm->broadcast((long long int *)id, 16, comm->root);
    if (b2g_comm_init(session->ctx, comm->size, comm->rank, id) != 0)
        throw std::runtime_error(std::string("b2g_comm_init: ") + b2g_last_error());
    mpo->tf = make_shared<GPUTensorFunctions<S, ParallelTensorFunctions<S, double>>>(mpo->tf->opf, prule, session);
    return session;
}

} // namespace b2g_host

