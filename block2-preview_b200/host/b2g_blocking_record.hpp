// b2g_blocking_record.hpp — reference-side recording of the blocking step (left_contract /
// right_contract) in the compact term form of include/b2g.h (b2g_tp_term).  Compiled together with
// block2's own headers; no CUDA, no library calls: it only produces descriptors.
#pragma once
#include "b2g.h"
#include "block2_core.hpp"
#include <functional>
#include <stdexcept>

namespace b2g_host {

using namespace block2;

// Term collector shared by an OperatorFunctions object and its per-thread copies.
struct TermCollector {
    bool active = false;
    vector<vector<b2g_tp_term>> per_thread;
    // H_eff diagonal: one recorder per walking thread, filled through the reference's own
    // BatchGEMMSeq::tensor_product_diagonal / three_tensor_product_diagonal (core/batch_gemm.hpp:1110-1135)
    bool diag_active = false;
    vector<shared_ptr<BatchGEMMSeq<double>>> diag_seqs;
    TermCollector() : per_thread(max(1, threading->n_threads_global)) {}
    void begin_diag() {
        diag_seqs.assign((size_t)max(1, threading->n_threads_global), nullptr);
        for (auto &q : diag_seqs)
            q = make_shared<BatchGEMMSeq<double>>(0, SeqTypes::Auto);
        diag_active = true;
    }
    void end_diag() {
        diag_active = false;
        diag_seqs.clear();
    }
    // intermediates / numerical_transform: the iadd calls of the stock walk (core/tensor_functions.hpp:2404-2517),
    // recorded per walking thread the same way
    bool iadd_active = false;
    vector<shared_ptr<BatchGEMMSeq<double>>> iadd_seqs; // additions with a block factor != 1 (GEMM form)
    vector<vector<b2g_tp_term>> iadd_terms;             // c_block += f * op(b_block), one descriptor per block
    void begin_iadd() {
        const size_t nt = (size_t)max(1, threading->n_threads_global);
        iadd_seqs.assign(nt, nullptr);
        for (auto &q : iadd_seqs)
            q = make_shared<BatchGEMMSeq<double>>(0, SeqTypes::Auto);
        iadd_terms.assign(nt, vector<b2g_tp_term>());
        iadd_active = true;
    }
    void end_iadd() {
        iadd_active = false;
        iadd_seqs.clear(), iadd_terms.clear();
    }
    void clear() {
        for (auto &v : per_thread)
            v.clear();
    }
    size_t size() const {
        size_t n = 0;
        for (auto &v : per_thread)
            n += v.size();
        return n;
    }
};

// OperatorFunctions whose tensor_product (core/operator_functions.hpp:672-711) can emit, instead of the
// per-row GEMM groups of AdvancedGEMM::tensor_product, one b2g_tp_term per connection-info entry - the
// arguments of the eager GMatrixFunctions::tensor_product call the stock method would make
// (core/matrix_functions.hpp:1269).  OperatorFunctions subclasses dispatch on every thread of
// TensorFunctions::parallel_for (the opf pointer survives the base-class slice), unlike
// fine-grained TensorFunctions overrides.
template <typename S> struct GPUOperatorFunctions : OperatorFunctions<S, double> {
    typedef double FL;
    typedef OperatorFunctions<S, double> Base;
    using Base::cg;
    using Base::seq;
    shared_ptr<TermCollector> collector;
    GPUOperatorFunctions(const shared_ptr<CG<S>> &cg, const shared_ptr<TermCollector> &collector)
        : Base(cg), collector(collector) {}
    shared_ptr<OperatorFunctions<S, FL>> copy() const override {
        shared_ptr<GPUOperatorFunctions<S>> r = make_shared<GPUOperatorFunctions<S>>(cg, collector);
        r->seq = seq->copy();
        return r;
    }
    // While the H_eff diagonal is recorded, the enumeration of the stock methods (operator_functions.hpp:
    // 211-328) runs unchanged but against this thread's own recorder, so that nothing lands in the shared
    // sequence TensorFunctions::tensor_product_diagonal executes on the host when the walk is over.
    struct SeqSwap {
        shared_ptr<BatchGEMMSeq<FL>> &slot, saved;
        SeqSwap(shared_ptr<BatchGEMMSeq<FL>> &slot, const shared_ptr<BatchGEMMSeq<FL>> &tmp) : slot(slot), saved(slot) {
            slot = tmp;
        }
        ~SeqSwap() { slot = saved; }
    };
    void tensor_product_diagonal(uint8_t conj, const shared_ptr<SparseMatrix<S, FL>> &a,
                                 const shared_ptr<SparseMatrix<S, FL>> &b, const shared_ptr<SparseMatrix<S, FL>> &c, S opdq,
                                 FL scale = 1.0) const override {
        if (!collector->diag_active)
            return Base::tensor_product_diagonal(conj, a, b, c, opdq, scale);
        SeqSwap sw(const_cast<GPUOperatorFunctions *>(this)->seq, collector->diag_seqs.at(threading->get_thread_id()));
        Base::tensor_product_diagonal(conj, a, b, c, opdq, scale);
    }
    void three_tensor_product_diagonal(uint8_t conj, const shared_ptr<SparseMatrix<S, FL>> &a,
                                       const shared_ptr<SparseMatrix<S, FL>> &b, const shared_ptr<SparseMatrix<S, FL>> &c,
                                       uint8_t dconj, const shared_ptr<SparseMatrix<S, FL>> &da,
                                       const shared_ptr<SparseMatrix<S, FL>> &db, bool dleft, S opdq,
                                       FL scale = 1.0) const override {
        if (!collector->diag_active)
            return Base::three_tensor_product_diagonal(conj, a, b, c, dconj, da, db, dleft, opdq, scale);
        SeqSwap sw(const_cast<GPUOperatorFunctions *>(this)->seq, collector->diag_seqs.at(threading->get_thread_id()));
        Base::three_tensor_product_diagonal(conj, a, b, c, dconj, da, db, dleft, opdq, scale);
    }
    void iadd(const shared_ptr<SparseMatrix<S, FL>> &a, const shared_ptr<SparseMatrix<S, FL>> &b, FL scale = 1.0,
              bool conj = false) const override {
        if (!collector->iadd_active)
            return Base::iadd(a, b, scale, conj);
        const int tid = threading->get_thread_id();
        if (a->factor != (FL)1.0 || a->total_memory >= (size_t)INT32_MAX) { // rare: keep the recorded GEMM form
            SeqSwap sw(const_cast<GPUOperatorFunctions *>(this)->seq, collector->iadd_seqs.at(tid));
            return Base::iadd(a, b, scale, conj);
        }
        // One descriptor per sector block instead of the per-row GEMM groups of the recorder: the block walk of
        // the stock method (core/operator_functions.hpp:135-174), emitting c_block += f * op(b_block) as a
        // tensor-product term with the 1 x 1 unit block as second factor.
        if (abs(b->factor * scale) < TINY)
            return;
        static const double one = 1.0;
        vector<b2g_tp_term> &out = collector->iadd_terms.at(tid);
        b2g_tp_term t;
        t.b = &one, t.bm = t.bn = 1, t.conjb = 0, t.reserved = 0;
        if (a->info == b->info && !conj) {
            t.a = b->data, t.c = a->data, t.am = 1, t.an = (int32_t)a->total_memory, t.cn = (int32_t)a->total_memory;
            t.conja = 0, t.scale = scale * b->factor;
            out.push_back(t);
            return;
        }
        const S bdq = b->info->delta_quantum;
        for (int ia = 0, ib; ia < a->info->n; ia++) {
            const S bra = a->info->quanta[ia].get_bra(a->info->delta_quantum), ket = a->info->quanta[ia].get_ket();
            const S bq = conj ? bdq.combine(ket, bra) : bdq.combine(bra, ket);
            if (bq == S(S::invalid) || (ib = b->info->find_state(bq)) == -1)
                continue;
            FL factor = scale * b->factor;
            if (conj)
                factor *= cg->transpose_cg(bdq, bra, ket);
            const GMatrix<FL> ma = (*a)[ia], mb = (*b)[ib];
            t.a = mb.data, t.c = ma.data, t.am = mb.m, t.an = mb.n, t.cn = ma.n, t.conja = conj ? 1 : 0, t.scale = factor;
            out.push_back(t);
        }
    }
    void tensor_product(uint8_t conj, const shared_ptr<SparseMatrix<S, FL>> &a, const shared_ptr<SparseMatrix<S, FL>> &b,
                        const shared_ptr<SparseMatrix<S, FL>> &c, FL scale = 1.0) const override {
        if (!collector->active)
            return Base::tensor_product(conj, a, b, c, scale);
        scale = scale * a->factor * b->factor;
        if (abs(scale) < TINY)
            return;
        const S adq = a->info->delta_quantum, bdq = b->info->delta_quantum, cdq = c->info->delta_quantum;
        const auto &ci = c->info->cinfo;
        // the (conj, a (x) b quantum) range of the connection table, as the stock method finds it
        const S abdq = cdq.combine((conj & 1) ? -adq : adq, (conj & 2) ? bdq : -bdq);
        const int ik = (int)(lower_bound(ci->quanta + ci->n[conj], ci->quanta + ci->n[conj + 1], abdq) - ci->quanta);
        assert(ik < ci->n[conj + 1]);
        const int lo = ci->idx[ik], hi = ik == ci->n[4] - 1 ? ci->nc : ci->idx[ik + 1];
        vector<b2g_tp_term> &out = collector->per_thread[threading->get_thread_id()];
        for (int il = lo; il < hi; il++) {
            const GMatrix<FL> ma = (*a)[ci->ia[il]], mb = (*b)[ci->ib[il]], mc = (*c)[ci->ic[il]];
            b2g_tp_term t;
            t.a = ma.data, t.b = mb.data, t.c = mc.data + ci->stride[il];
            t.am = ma.m, t.an = ma.n, t.bm = mb.m, t.bn = mb.n, t.cn = mc.n;
            t.conja = conj & 1, t.conjb = (conj & 2) >> 1, t.reserved = 0;
            t.scale = scale * (FL)ci->factor[il];
            out.push_back(t);
        }
    }
};

// Record-only walk of one blocking expression, the Auto-mode counterpart of
// TensorFunctions::tensor_product (core/tensor_functions.hpp:2184-2288).  Products go to the
// recorder of `opf` through the reference's own OperatorFunctions::tensor_product.  A SumProd
// term whose pre-sum is not stored as an intermediate needs a temporary tmp = sum_i f_i op_i
// BEFORE the product that reads it: the stock method frees tmp right after recording, which is
// only valid when the list is executed at record time (Simple), so here the iadd entries go to
// a second recorder (`pre`, executed first) and the temporaries stay alive in `temps` until
// both lists have run.
template <typename S>
inline void record_blocking_expr(const shared_ptr<OperatorFunctions<S, double>> &opf, const shared_ptr<OpExpr<S>> &expr,
                                 const unordered_map<shared_ptr<OpExpr<S>>, shared_ptr<SparseMatrix<S, double>>> &lop,
                                 const unordered_map<shared_ptr<OpExpr<S>>, shared_ptr<SparseMatrix<S, double>>> &rop,
                                 const shared_ptr<SparseMatrix<S, double>> &mat,
                                 const shared_ptr<OperatorFunctions<S, double>> &pre,
                                 vector<shared_ptr<SparseMatrix<S, double>>> &temps,
                                 const std::function<void(const shared_ptr<SparseMatrix<S, double>> &,
                                                          const shared_ptr<SparseMatrixInfo<S>> &)> *alloc_tmp = nullptr) {
    typedef double FL;
    typedef unordered_map<shared_ptr<OpExpr<S>>, shared_ptr<SparseMatrix<S, FL>>> OpMap;
    const OpTypes ty = expr->get_type();
    if (ty == OpTypes::Zero)
        return;
    if (ty == OpTypes::Sum) {
        for (auto &x : dynamic_pointer_cast<OpSum<S, FL>>(expr)->strings)
            record_blocking_expr<S>(opf, x->get_type() == OpTypes::Prod && x->b == nullptr
                                     ? (shared_ptr<OpExpr<S>>)x->get_op()
                                     : (shared_ptr<OpExpr<S>>)x,
                                 lop, rop, mat, pre, temps, alloc_tmp);
        return;
    }
    if (ty == OpTypes::Elem) { // singlet embedding: the partner is the identity of the other block
        auto op = dynamic_pointer_cast<OpElement<S, FL>>(expr);
        const shared_ptr<OpExpr<S>> ident = make_shared<OpExpr<S>>();
        opf->tensor_product(0, lop.count(op) ? lop.at(op) : lop.at(ident), rop.count(op) ? rop.at(op) : rop.at(ident),
                            mat, op->factor);
        return;
    }
    if (ty == OpTypes::Prod) {
        auto op = dynamic_pointer_cast<OpProduct<S, FL>>(expr);
        opf->tensor_product(op->conj, lop.at(op->a), rop.at(op->b), mat, op->factor);
        return;
    }
    if (ty != OpTypes::SumProd)
        throw std::runtime_error("b2g: unexpected expression type in a blocking expression");
    auto op = dynamic_pointer_cast<OpSumProd<S, FL>>(expr);
    const bool sum_right = op->b == nullptr; // a (x) (sum of right operators), else (sum of left) (x) b
    const OpMap &side = sum_right ? rop : lop;
    shared_ptr<SparseMatrix<S, FL>> sum;
    if (op->c != nullptr && side.count(op->c))
        sum = side.at(op->c); // stored intermediate
    else {
        sum = make_shared<SparseMatrix<S, FL>>(make_shared<VectorAllocator<FL>>());
        const shared_ptr<SparseMatrixInfo<S>> &sinfo = side.at(abs_value((shared_ptr<OpExpr<S>>)op->ops[0]))->info;
        if (alloc_tmp != nullptr) // storage from the caller (reserved addresses, structure-only arenas)
            (*alloc_tmp)(sum, sinfo);
        else
            sum->allocate(sinfo);
        if (pre != nullptr) // structure-only recording passes no pre-sum recorder
            for (size_t i = 0; i < op->ops.size(); i++)
                pre->iadd(sum, side.at(abs_value((shared_ptr<OpExpr<S>>)op->ops[i])), op->ops[i]->factor,
                          op->conjs[i]);
        temps.push_back(sum);
    }
    if (sum_right)
        opf->tensor_product(op->conj, lop.at(op->a), sum, mat, op->factor);
    else
        opf->tensor_product(op->conj, sum, rop.at(op->b), mat, op->factor);
}

} // namespace b2g_host
