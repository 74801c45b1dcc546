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
#include "b2g_shm_comm.hpp"
#include <atomic>
#include <stdexcept>

namespace b2g_host {

using namespace block2;

struct Session {
    b2g_context *ctx = nullptr;
    std::atomic<bool> recording{false};
    double t_plan = 0, t_matvec = 0;
    size_t n_plan = 0, n_matvec = 0;
    // --verify: worst relative deviation ||sigma_gpu - sigma_cpu|| / ||sigma_cpu|| over all sites,
    // sigma_cpu from the reference's own BatchGEMMSeq::operator() on the same recorded list
    bool verify = false;
    double max_matvec_err = 0;
    size_t n_verified = 0;
    // renormalisation (left_rotate / right_rotate) on the device
    bool gpu_rotate = false;
    double t_rotate = 0, max_rotate_err = 0;
    size_t n_rotate = 0, rotate_pairs = 0;
    double rotate_flops = 0;
    // blocking (left_contract / right_contract) on the device
    bool gpu_contract = false;
    double t_contract = 0, max_contract_err = 0, contract_kernel_ms = 0, contract_bytes = 0;
    double t_contract_record = 0, t_contract_plan = 0, t_contract_upload = 0, t_contract_download = 0;
    // blocks the latest blocking calls left resident in HBM (b2g KEEP_RESIDENT): the operator, the host
    // address and size it had when it was produced.  An entry is vouched for (b2g_resident_vouch) only
    // while the operator object is alive and still owns exactly that storage.
    bool keep_resident = true, uninit_outputs = true;
    struct ResidentOp {
        std::weak_ptr<void> owner;
        const double *data;
        size_t doubles;
        const double *const *slot; // &SparseMatrix::data of the owner
        const size_t *size_slot;   // &SparseMatrix::total_memory
    };
    std::vector<ResidentOp> resident_ops;
    void vouch_residents() {
        std::vector<const double *> ptrs;
        std::vector<int64_t> sizes;
        for (auto &r : resident_ops) {
            std::shared_ptr<void> alive = r.owner.lock();
            if (alive != nullptr && *r.slot == r.data && *r.size_slot == r.doubles)
                ptrs.push_back(r.data), sizes.push_back((int64_t)r.doubles);
        }
        if (!ptrs.empty() && b2g_resident_vouch(ctx, (int64_t)ptrs.size(), ptrs.data(), sizes.data()) != 0)
            throw std::runtime_error(std::string("b2g_resident_vouch: ") + b2g_last_error());
    }
    void drop_residents() {
        resident_ops.clear();
        int64_t held = 0, hit = 0;
        b2g_resident_stats(ctx, &held, &hit);
        resident_hit_bytes = (double)hit, resident_peak_bytes = std::max(resident_peak_bytes, (double)held);
        b2g_resident_drop(ctx);
    }
    double resident_hit_bytes = 0, resident_peak_bytes = 0;
    size_t n_contract = 0, contract_entries = 0;
    explicit Session(int device = 0) {
        uninit_outputs = getenv("B2G_ZERO_OUTPUTS") == nullptr; // A/B switch, see contract_on_device
        if (b2g_context_create(device, &ctx) != 0)
            throw std::runtime_error(std::string("b2g_context_create: ") + b2g_last_error());
    }
    ~Session() { b2g_context_destroy(ctx); }
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

// Uninitialised storage for blocked operators whose every element the device result overwrites
// (B2G_DST_COVERED): VectorAllocator + SparseMatrix::allocate would zero the block twice first.
struct UninitAllocator : Allocator<double> {
    double *allocate(size_t n) override {
        double *p = (double *)malloc(std::max<size_t>(n, 1) * sizeof(double));
        if (p == nullptr)
            throw std::bad_alloc();
        return p;
    }
    void deallocate(void *ptr, size_t) override { free(ptr); }
    double *reallocate(double *ptr, size_t, size_t new_n) override {
        return (double *)realloc(ptr, std::max<size_t>(new_n, 1) * sizeof(double));
    }
};

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
            }
        }
    }
    // Renormalisation c = bra^T . a . ket of every operator of a block (core/tensor_functions.hpp:
    // 2365-2403).  The reference's own OperatorFunctions::tensor_rotate enumerates the sector blocks
    // and records one rotate() pair per block (operator_functions.hpp:175-210); in Auto mode nothing
    // is executed at record time, so the list is handed to b2g_pairs_execute instead of
    // seq->auto_perform().  c is zero-initialised by allocate() exactly as in the stock method.
    void rotate_on_device(const shared_ptr<OperatorTensor<S, FL>> &a, const shared_ptr<SparseMatrix<S, FL>> &mpst_bra,
                          const shared_ptr<SparseMatrix<S, FL>> &mpst_ket, shared_ptr<OperatorTensor<S, FL>> &c,
                          const shared_ptr<Symbolic<S>> &names, bool trans) const {
        Timer t;
        t.get_time();
        for (auto &p : c->ops)
            p.second->allocate(p.second->info);
        auto &seq = opf->seq;
        if (seq->batch[0]->gp.size() != 0 || seq->batch[1]->gp.size() != 0)
            throw std::runtime_error("b2g: recorder not empty at rotate");
        const SeqTypes saved = seq->mode;
        seq->mode = SeqTypes::Auto; // record only
        for (size_t i = 0; i < names->data.size(); i++)
            if (names->data[i]->get_type() != OpTypes::Zero) {
                auto pa = abs_value(names->data[i]);
                opf->tensor_rotate(a->ops.at(pa), c->ops.at(pa), mpst_bra, mpst_ket, trans);
            }
        seq->mode = saved;
        if (seq->batch[1]->gp.size() != 0) {
            b2g_batch b0 = as_b2g_batch(*seq->batch[0]), b1 = as_b2g_batch(*seq->batch[1]);
            session->vouch_residents(); // blocks the blocking step just produced are read from HBM
            if (session->verify) { // keep a CPU copy of the result to compare with
                vector<vector<double>> ref;
                for (auto &p : c->ops)
                    ref.emplace_back(p.second->data, p.second->data + p.second->total_memory);
                if (b2g_pairs_execute(session->ctx, &b0, &b1, 0, nullptr) != 0)
                    throw std::runtime_error(std::string("b2g_pairs_execute: ") + b2g_last_error());
                vector<vector<double>> gpu;
                size_t z = 0;
                for (auto &p : c->ops) {
                    gpu.emplace_back(p.second->data, p.second->data + p.second->total_memory);
                    memcpy(p.second->data, ref[z++].data(), sizeof(double) * p.second->total_memory);
                }
                seq->mode = SeqTypes::Auto;
                seq->auto_perform(); // the reference executor on the same list
                seq->mode = saved;
                double num = 0, den = 0;
                z = 0;
                for (auto &p : c->ops) {
                    for (size_t j = 0; j < p.second->total_memory; j++)
                        num += (gpu[z][j] - p.second->data[j]) * (gpu[z][j] - p.second->data[j]),
                            den += p.second->data[j] * p.second->data[j];
                    memcpy(p.second->data, gpu[z++].data(), sizeof(double) * p.second->total_memory);
                }
                session->max_rotate_err = max(session->max_rotate_err, den > 0 ? sqrt(num / den) : sqrt(num));
            } else {
                b2g_plan_stats st;
                if (b2g_pairs_execute(session->ctx, &b0, &b1, 0, &st) != 0)
                    throw std::runtime_error(std::string("b2g_pairs_execute: ") + b2g_last_error());
                session->rotate_pairs += (size_t)st.pairs, session->rotate_flops += 2.0 * (double)st.nflop_mnk;
                seq->cumulative_nflop += (size_t)st.nflop_mnk;
            }
        }
        seq->clear();
        session->drop_residents(); // the blocked operators are dead after their renormalisation
        session->t_rotate += t.get_time(), session->n_rotate++;
    }
    typedef unordered_map<shared_ptr<OpExpr<S>>, shared_ptr<SparseMatrix<S, FL>>> OpMap;
    int run_blocking_list(BatchGEMM<FL> &bt, b2g_blocking_stats &st) const {
        static_assert(sizeof(CBLAS_TRANSPOSE) == sizeof(int32_t) && sizeof(MKL_INT) == sizeof(int32_t), "");
        return b2g_batch_execute(session->ctx, (int64_t)bt.gp.size(), (const int32_t *)bt.ta.data(),
                                 (const int32_t *)bt.tb.data(), bt.m.data(), bt.n.data(), bt.k.data(), bt.alpha.data(),
                                 bt.a.data(), bt.lda.data(), bt.b.data(), bt.ldb.data(), bt.beta.data(), bt.c.data(),
                                 bt.ldc.data(), bt.gp.data(), B2G_OPERANDS_HOST, B2G_DST_ZERO, &st);
    }
    // Blocking c[op] = sum a[x] (x) b[y] (core/tensor_functions.hpp:2842-2885 / 2941-2984): allocate the
    // non-delayed, not yet cached operators as the stock method does, walk every expression in record-only
    // mode, and execute on the device what the walk produced: b2g_tp_term descriptors when opf is a
    // GPUOperatorFunctions (b2g_tensor_product_execute), otherwise the recorded GEMM list
    // (b2g_batch_execute, in place of seq->auto_perform()).
    void contract_on_device(const shared_ptr<OperatorTensor<S, FL>> &a, const shared_ptr<OperatorTensor<S, FL>> &b,
                            shared_ptr<OperatorTensor<S, FL>> &c, const shared_ptr<Symbolic<S>> &exprs,
                            const shared_ptr<Symbolic<S>> &names, OpNamesSet delayed, bool right) const {
        Timer t, tr;
        t.get_time();
        auto &seq = opf->seq;
        if (seq->batch[0]->gp.size() != 0 || seq->batch[1]->gp.size() != 0)
            throw std::runtime_error("b2g: recorder not empty at contract");
        assert(exprs->data.size() == names->data.size());
        const SeqTypes saved = seq->mode;
        shared_ptr<GPUOperatorFunctions<S>> gopf = dynamic_pointer_cast<GPUOperatorFunctions<S>>(opf);
        const OpMap &lop = right ? b->ops : a->ops, &rop = right ? a->ops : b->ops;
        // uninitialised outputs + overwrite (B2G_DST_COVERED) instead of zero fill (twice: VectorAllocator and
        // SparseMatrix::allocate) + add: the first touch of the pages happens in the threaded, pipelined
        // download.  C2 M=1000, three sweeps: blocking 11.8 -> 5.6 s, sweeps 21.2 -> 16.9 s
        const bool covered = gopf != nullptr && session->keep_resident && session->uninit_outputs;
        vector<size_t> todo;
        for (size_t i = 0; i < exprs->data.size(); i++) {
            shared_ptr<OpElement<S, FL>> cop = dynamic_pointer_cast<OpElement<S, FL>>(names->data[i]);
            shared_ptr<SparseMatrix<S, FL>> &m = c->ops.at(abs_value(names->data[i]));
            if (delayed(cop->name) || m->alloc != nullptr) // delayed, or the cached part
                continue;
            if (covered) { // every element is overwritten by the device result: no zero fill
                m->alloc = make_shared<UninitAllocator>();
                const size_t n = m->info->template get_total_memory<FL>();
                m->allocate(m->info, n == 0 ? nullptr : m->alloc->allocate(n));
            } else {
                m->alloc = make_shared<VectorAllocator<FL>>();
                m->allocate(m->info);
            }
            todo.push_back(i);
        }
        // record-only walk; the SumProd pre-sums go to `pre`, their temporaries to `temps`
        auto walk = [&](const shared_ptr<OperatorFunctions<S, FL>> &pre, vector<shared_ptr<SparseMatrix<S, FL>>> &temps) {
            seq->mode = SeqTypes::Auto;
            pre->seq->mode = SeqTypes::Auto;
            for (size_t i : todo) {
                shared_ptr<OpElement<S, FL>> cop = dynamic_pointer_cast<OpElement<S, FL>>(names->data[i]);
                record_blocking_expr<S>(opf, exprs->data[i] * ((FL)1.0 / cop->factor), lop, rop,
                                        c->ops.at(abs_value(names->data[i])), pre, temps);
            }
            seq->mode = saved;
            if (seq->batch[0]->gp.size() != 0 || pre->seq->batch[0]->gp.size() != 0)
                throw std::runtime_error("b2g: blocking list has chained pairs");
        };
        auto account = [&](const b2g_blocking_stats &st) {
            session->contract_entries += (size_t)st.entries, session->contract_kernel_ms += st.kernel_ms;
            session->contract_bytes += (double)(st.bytes_in + st.bytes_out);
            session->t_contract_plan += st.plan_seconds, session->t_contract_upload += st.upload_seconds;
            session->t_contract_download += st.download_seconds;
            seq->cumulative_nflop += (size_t)st.nflop_mnk;
        };
        b2g_blocking_stats st;
        vector<shared_ptr<SparseMatrix<S, FL>>> temps;
        vector<shared_ptr<OperatorFunctions<S, FL>>> pres; // pre-sum recorders (one per recording thread)
        tr.get_time();
        if (gopf != nullptr) {
            // term form: the operators are walked by the operator-level threads, as the stock method does
            // with parallel_for; every thread has its own term vector, pre-sum recorder and temporaries,
            // and an operator is walked by one thread, so its terms stay in expression order
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
            std::exception_ptr err = nullptr;
#pragma omp parallel for schedule(dynamic) num_threads(nt)
            for (int z = 0; z < (int)todo.size(); z++) {
                const int tid = threading->get_thread_id();
                const size_t i = todo[z];
                try {
                    shared_ptr<OpElement<S, FL>> cop = dynamic_pointer_cast<OpElement<S, FL>>(names->data[i]);
                    record_blocking_expr<S>(opfs[tid], exprs->data[i] * ((FL)1.0 / cop->factor), lop, rop,
                                            c->ops.at(abs_value(names->data[i])), pres[tid], temps_t[tid]);
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
        } else {
            pres.push_back(make_shared<OperatorFunctions<S, FL>>(opf->cg));
            walk(pres[0], temps);
        }
        session->t_contract_record += tr.get_time();
        for (auto &pre : pres) {
            if (pre->seq->batch[0]->gp.size() != 0)
                throw std::runtime_error("b2g: blocking list has chained pairs");
            if (pre->seq->batch[1]->gp.size() != 0) {
                if (run_blocking_list(*pre->seq->batch[1], st) != 0)
                    throw std::runtime_error(std::string("b2g_batch_execute: ") + b2g_last_error());
                account(st);
            }
        }
        if (gopf != nullptr) {
            vector<b2g_tp_term> &terms = gopf->collector->per_thread[0];
            for (size_t k = 1; k < gopf->collector->per_thread.size(); k++)
                terms.insert(terms.end(), gopf->collector->per_thread[k].begin(), gopf->collector->per_thread[k].end());
            if (terms.size() != 0) {
                if (session->keep_resident) { // the resident mirror covers the whole fresh operators
                    vector<const double *> cp;
                    vector<int64_t> cn;
                    for (size_t i : todo) {
                        auto &m = c->ops.at(abs_value(names->data[i]));
                        cp.push_back(m->data), cn.push_back((int64_t)m->total_memory);
                    }
                    b2g_resident_cover(session->ctx, (int64_t)cp.size(), cp.data(), cn.data());
                }
                if (b2g_tensor_product_execute(session->ctx, (int64_t)terms.size(), terms.data(), B2G_OPERANDS_HOST,
                                               B2G_DST_ZERO | (session->keep_resident ? B2G_KEEP_RESIDENT : 0) |
                                                   (covered ? B2G_DST_COVERED : 0),
                                               &st) != 0)
                    throw std::runtime_error(std::string("b2g_tensor_product_execute: ") + b2g_last_error());
                account(st);
            } else if (covered) // nothing writes the fresh operators: they are zero
                for (size_t i : todo) {
                    auto &m = c->ops.at(abs_value(names->data[i]));
                    if (m->total_memory != 0)
                        memset(m->data, 0, sizeof(double) * m->total_memory);
                }
            gopf->collector->clear();
            if (session->keep_resident)
                for (size_t i : todo) {
                    shared_ptr<SparseMatrix<S, FL>> m = c->ops.at(abs_value(names->data[i]));
                    session->resident_ops.push_back(typename Session::ResidentOp{
                        std::weak_ptr<void>(std::shared_ptr<void>(m)), m->data, (size_t)m->total_memory,
                        (const double *const *)&m->data, (const size_t *)&m->total_memory});
                }
        } else if (seq->batch[1]->gp.size() != 0) {
            if (run_blocking_list(*seq->batch[1], st) != 0)
                throw std::runtime_error(std::string("b2g_batch_execute: ") + b2g_last_error());
            account(st);
        }
        seq->clear();
        for (auto &pre : pres)
            pre->seq->clear();
        if (session->verify && todo.size() != 0) {
            // the reference's own executor on the list its own recorder makes of the same expressions
            vector<vector<double>> gpu;
            for (size_t i : todo) {
                auto &m = c->ops.at(abs_value(names->data[i]));
                gpu.emplace_back(m->data, m->data + m->total_memory);
                memset(m->data, 0, sizeof(double) * m->total_memory);
            }
            vector<shared_ptr<SparseMatrix<S, FL>>> temps2;
            shared_ptr<OperatorFunctions<S, FL>> pre2 = make_shared<OperatorFunctions<S, FL>>(opf->cg);
            walk(pre2, temps2);
            pre2->seq->auto_perform();
            seq->mode = SeqTypes::Auto;
            seq->auto_perform();
            seq->mode = saved;
            seq->clear(), pre2->seq->clear();
            double num = 0, den = 0;
            size_t z = 0;
            for (size_t i : todo) {
                auto &m = c->ops.at(abs_value(names->data[i]));
                for (size_t j = 0; j < m->total_memory; j++)
                    num += (gpu[z][j] - m->data[j]) * (gpu[z][j] - m->data[j]), den += m->data[j] * m->data[j];
                memcpy(m->data, gpu[z++].data(), sizeof(double) * m->total_memory);
            }
            session->max_contract_err = max(session->max_contract_err, den > 0 ? sqrt(num / den) : sqrt(num));
            for (auto &m : temps2)
                m->deallocate();
        }
        for (auto &m : temps)
            m->deallocate();
        session->t_contract += t.get_time(), session->n_contract++;
    }
    void left_contract(const shared_ptr<OperatorTensor<S, FL>> &a, const shared_ptr<OperatorTensor<S, FL>> &b,
                       shared_ptr<OperatorTensor<S, FL>> &c, const shared_ptr<Symbolic<S>> &cexprs = nullptr,
                       OpNamesSet delayed = OpNamesSet()) const override {
        if (!session->gpu_contract || session->recording || a == nullptr || prule != nullptr ||
            frame_<FL>()->use_main_stack)
            return Base::left_contract(a, b, c, cexprs, delayed);
        contract_on_device(a, b, c, cexprs == nullptr ? a->lmat * b->lmat : cexprs, c->lmat, delayed, false);
    }
    void right_contract(const shared_ptr<OperatorTensor<S, FL>> &a, const shared_ptr<OperatorTensor<S, FL>> &b,
                        shared_ptr<OperatorTensor<S, FL>> &c, const shared_ptr<Symbolic<S>> &cexprs = nullptr,
                        OpNamesSet delayed = OpNamesSet()) const override {
        if (!session->gpu_contract || session->recording || a == nullptr || prule != nullptr ||
            frame_<FL>()->use_main_stack)
            return Base::right_contract(a, b, c, cexprs, delayed);
        contract_on_device(a, b, c, cexprs == nullptr ? b->rmat * a->rmat : cexprs, c->rmat, delayed, true);
    }
    void left_rotate(const shared_ptr<OperatorTensor<S, FL>> &a, const shared_ptr<SparseMatrix<S, FL>> &mpst_bra,
                     const shared_ptr<SparseMatrix<S, FL>> &mpst_ket,
                     shared_ptr<OperatorTensor<S, FL>> &c) const override {
        if (!session->gpu_rotate || session->recording)
            return Base::left_rotate(a, mpst_bra, mpst_ket, c);
        rotate_on_device(a, mpst_bra, mpst_ket, c, a->lmat, false);
    }
    void right_rotate(const shared_ptr<OperatorTensor<S, FL>> &a, const shared_ptr<SparseMatrix<S, FL>> &mpst_bra,
                      const shared_ptr<SparseMatrix<S, FL>> &mpst_ket,
                      shared_ptr<OperatorTensor<S, FL>> &c) const override {
        if (!session->gpu_rotate || session->recording)
            return Base::right_rotate(a, mpst_bra, mpst_ket, c);
        rotate_on_device(a, mpst_bra, mpst_ket, c, a->rmat, true);
    }
    void build_plan() const {
        Timer t;
        t.get_time();
        drop();
        auto &seq = opf->seq;
        b2g_batch b0 = as_b2g_batch(*seq->batch[0]), b1 = as_b2g_batch(*seq->batch[1]);
        session->vouch_residents(); // complementary operators blocked for this H_eff are still in HBM
        if (b2g_plan_create(session->ctx, &b0, &b1, (int64_t)seq->max_work, (int64_t)csize, (int64_t)vsize,
                            B2G_OPERANDS_HOST, &plan) != 0)
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
        if (opf->seq->batch[0]->gp.size() == 0)
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
        if (!plain)
            return Base::two_dot_eigs_and_perturb(forward, i, davidson_conv_thrd, noise, pket);
        Timer t;
        t.get_time();
        shared_ptr<EffectiveHamiltonian<S, double>> h_eff =
            me->eff_ham(FuseTypes::FuseLR, forward, true, me->bra->tensors[i], me->ket->tensors[i]);
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
        h_eff->precompute();
        if (gtf->session->verify && h_eff->tf->opf->seq->batch[0]->gp.size() != 0) {
            const size_t n = h_eff->ket->total_memory;
            vector<double> x(n), y_gpu(n, 0.0), y_cpu(n, 0.0);
            Random::fill<double>(x.data(), n);
            GMatrix<double> xm(x.data(), (MKL_INT)n, 1);
            (*gtf)(xm, GMatrix<double>(y_gpu.data(), (MKL_INT)n, 1), 1.0);
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
        if (h_eff->tf->opf->seq->batch[0]->gp.size() != 0) {
            if (b2g_davidson(gtf->get_plan(), h_eff->diag->data, h_eff->ket->data, davidson_conv_thrd,
                             this->davidson_rel_conv_thrd, this->davidson_max_iter, this->davidson_soft_max_iter,
                             this->davidson_def_min_size, this->davidson_def_max_size, &e, &ndav) != 0)
                throw std::runtime_error(std::string("b2g_davidson: ") + b2g_last_error());
            nflop = (size_t)ndav * (h_eff->tf->opf->seq->batch[0]->nflop + h_eff->tf->opf->seq->batch[1]->nflop);
        }
        h_eff->post_precompute();
        gtf->drop();
        double tdav = t.get_time();
        this->teig += tdav;
        h_eff->deallocate();
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
    if (pmpo == nullptr)
        throw std::runtime_error("b2g: install_parallel needs a ParallelMPO");
    shared_ptr<Session> session = make_shared<Session>(device);
    auto comm = pmpo->rule->comm;
    char id[128];
    if (comm->rank == comm->root && b2g_comm_unique_id(id) != 0)
        throw std::runtime_error(std::string("b2g_comm_unique_id: ") + b2g_last_error());
    static_assert(sizeof(long long int) == 8, "");
    comm->broadcast((long long int *)id, 16, comm->root);
    if (b2g_comm_init(session->ctx, comm->size, comm->rank, id) != 0)
        throw std::runtime_error(std::string("b2g_comm_init: ") + b2g_last_error());
    mpo->tf = make_shared<GPUTensorFunctions<S, ParallelTensorFunctions<S, double>>>(mpo->tf->opf, pmpo->rule, session);
    return session;
}

} // namespace b2g_host
