// b2g_dmrg.cpp — host driver: block2's own two-site DMRG sweep with the H.C matvec (and,
// optionally, the whole Davidson solver) executed by libb2g.so on the GPU.
//
// Built against the reference headers where they lie (-I$REF/src), never copied.  The only
// lines that differ from a stock block2 driver are `b2g_host::install(mpo)` and the choice of
// b2g_host::GPUDMRG instead of DMRG.  With --compare the stock CPU path runs first on the same
// FCIDUMP, seed and sweep schedule, and the per-sweep energy differences are printed
// (north-star bar: 1e-8 Ha at equal bond dimension).
#include "b2g_adapter.hpp"
#include <cstdio>
#include <cstring>

using namespace block2;
using namespace std;

#ifndef B2G_S
#define B2G_S SU2
#endif

// ---- LAPACK hook of the density-matrix split (b2g_adapter.hpp: SplitHook); one definition per binary
extern "C" void scipy_dsyev_(const char *jobz, const char *uplo, const MKL_INT *n, double *a, const MKL_INT *lda,
                             double *w, double *work, const MKL_INT *lwork, MKL_INT *info);
// the symbol the reference's dsyev_ calls resolve to in this build (blas_rename.h)
extern "C" void b2g_host_dsyev_(const char *jobz, const char *uplo, const MKL_INT *n, double *a,
                                       const MKL_INT *lda, double *w, double *work, const MKL_INT *lwork,
                                       MKL_INT *info) {
    b2g_context *ctx = b2g_host::SplitHook::ctx();
    if (ctx != nullptr && *lwork != -1 && *n >= b2g_host::SplitHook::min_n() && (jobz[0] == 'V' || jobz[0] == 'v') &&
        b2g_syevd(ctx, (int)*n, a, (int)*lda, w) == 0) {
        *info = 0;
        b2g_host::SplitHook::calls_gpu()++;
        return;
    }
    if (ctx != nullptr && *lwork != -1)
        b2g_host::SplitHook::calls_cpu()++;
    scipy_dsyev_(jobz, uplo, n, a, lda, w, work, lwork, info);
}

struct Args {
    string fcidump, pg = "d2h", occ = "", scratch = "/tmp/b2g_scratch", davidson = "device";
    int bond = 250, n_sweeps = 6, threads = 8, device = 0, ranks = 1, rank = 0;
    // never 0: Random::rand_seed(0) seeds from the clock (core/utils.hpp:231-236) and the two arms of
    // --compare would start from different MPS
    int seed = 1234;
    string shm = "b2g";
    bool compare = false, verify = false, gpu_rotate = true, gpu_contract = true, gpu_diag = true, gpu_iadd = true,
         host_mirror = false, pin = true, cpu_only = false, classic = false, gpu_split = false;
    int restart_sweeps = 0; // --compare: zero-noise sweeps both arms run from the CPU arm's final MPS (same state)
    int noise_sweeps = 2;   // sweeps per noise level: {noise x k, 0.1 noise x k, 0 ...}
    double dav_thrd = 0;    // > 0: Davidson threshold of every sweep (default: the reference's noise-derived schedule)
    double conv = 1e-7, noise = 1e-5;
    size_t dsize_gb = 8;
};

static PGTypes pg_of(const string &s) {
    if (s == "d2h") return PGTypes::D2H;
    if (s == "c2v") return PGTypes::C2V;
    if (s == "c2h") return PGTypes::C2H;
    if (s == "d2") return PGTypes::D2;
    if (s == "cs") return PGTypes::CS;
    if (s == "c2") return PGTypes::C2;
    if (s == "ci") return PGTypes::CI;
    return PGTypes::C1;
}

struct RunResult {
    vector<double> energies;
    double total = 0, teig = 0, teff = 0, tblk = 0;
    size_t nflop = 0;
};

template <typename S>
static RunResult run_dmrg(const Args &args, const shared_ptr<MPO<S, double>> &mpo,
                          const shared_ptr<HamiltonianQC<S, double>> &hamil, S target, bool gpu) {
    ubond_t bond_dim = (ubond_t)args.bond;
    shared_ptr<MPSInfo<S>> mps_info = make_shared<MPSInfo<S>>(hamil->n_sites, hamil->vacuum, target, hamil->basis);
    if (args.occ != "")
        mps_info->set_bond_dimension_using_occ(bond_dim, read_occ(args.occ), 1);
    else
        mps_info->set_bond_dimension(bond_dim);
    Random::rand_seed(args.seed);
    shared_ptr<MPS<S, double>> mps = make_shared<MPS<S, double>>(hamil->n_sites, 0, 2);
    mps->initialize(mps_info);
    mps->random_canonicalize();
    mps->save_mutable();
    mps->deallocate();
    mps_info->save_mutable();
    mps_info->deallocate_mutable();
    shared_ptr<MovingEnvironment<S, double, double>> me =
        make_shared<MovingEnvironment<S, double, double>>(mpo, mps, mps, "DMRG");
    me->init_environments(false);
    me->delayed_contraction = OpNamesSet::normal_ops();
    me->cached_contraction = true;
    vector<ubond_t> bdims = {bond_dim};
    vector<double> noises;
    for (int z = 0; z < 2 * args.noise_sweeps; z++)
        noises.push_back(z < args.noise_sweeps ? args.noise : args.noise * 0.1);
    noises.push_back(0.0);
    if (args.noise == 0)
        noises = {0.0};
    shared_ptr<DMRG<S, double, double>> dmrg;
    if (gpu) {
        auto g = make_shared<b2g_host::GPUDMRG<S>>(me, bdims, noises);
        g->device_davidson = args.davidson == "device";
        dmrg = g;
    } else
        dmrg = make_shared<DMRG<S, double, double>>(me, bdims, noises);
    dmrg->iprint = 2;
    dmrg->noise_type = NoiseTypes::DensityMatrix;
    dmrg->decomp_type = DecompositionTypes::DensityMatrix;
    dmrg->davidson_soft_max_iter = 4000;
    if (args.dav_thrd > 0)
        dmrg->davidson_conv_thrds = vector<double>(noises.size(), args.dav_thrd);
    Timer t;
    t.get_time();
    dmrg->solve(args.n_sweeps, true, args.conv * 0.1);
    RunResult r;
    r.total = t.get_time();
    for (auto &e : dmrg->energies)
        r.energies.push_back((double)e[0]);
    if (!gpu && args.restart_sweeps > 0) {
        // keep the converged state of the CPU arm on disk under its own tag: both arms continue from it
        shared_ptr<MPS<S, double>> cp = mps->deep_copy("B2GEND");
        cp->info->save_data(frame_<double>()->mps_dir + "/B2GEND-mps_info.bin");
        cp->info->deallocate();
    }
    mps_info->deallocate();
    me->remove_partition_files();
    return r;
}

// Zero-noise sweeps from the state the CPU arm ended in (tag B2GEND), on a private copy: the two arms start from
// the SAME MPS, so the energy of every sweep is comparable to rounding (no random initial state, no noise draws).
template <typename S>
static vector<double> continue_dmrg(const Args &args, const shared_ptr<MPO<S, double>> &mpo, bool gpu) {
    shared_ptr<MPSInfo<S>> info = make_shared<MPSInfo<S>>(0);
    info->load_data(frame_<double>()->mps_dir + "/B2GEND-mps_info.bin");
    info->load_mutable();
    shared_ptr<MPS<S, double>> mps = make_shared<MPS<S, double>>(info);
    mps->load_data();
    mps->load_mutable();
    info->tag = gpu ? "B2GCG" : "B2GCR";
    info->save_mutable();
    mps->save_mutable();
    mps->save_data();
    mps->deallocate();
    info->deallocate_mutable();
    shared_ptr<MovingEnvironment<S, double, double>> me =
        make_shared<MovingEnvironment<S, double, double>>(mpo, mps, mps, "DMRG");
    me->init_environments(false);
    me->delayed_contraction = OpNamesSet::normal_ops();
    me->cached_contraction = true;
    vector<ubond_t> bdims = {(ubond_t)args.bond};
    vector<double> noises = {0.0};
    shared_ptr<DMRG<S, double, double>> dmrg;
    if (gpu)
        dmrg = make_shared<b2g_host::GPUDMRG<S>>(me, bdims, noises);
    else
        dmrg = make_shared<DMRG<S, double, double>>(me, bdims, noises);
    dmrg->iprint = 2;
    dmrg->noise_type = NoiseTypes::DensityMatrix;
    dmrg->decomp_type = DecompositionTypes::DensityMatrix;
    dmrg->davidson_soft_max_iter = 4000;
    dmrg->davidson_conv_thrds = vector<double>(1, args.dav_thrd > 0 ? args.dav_thrd : 1e-10);
    dmrg->solve(args.restart_sweeps, mps->center == 0, 0.0);
    vector<double> e;
    for (auto &x : dmrg->energies)
        e.push_back((double)x[0]);
    info->deallocate();
    me->remove_partition_files();
    return e;
}

int main(int argc, char **argv) {
    Args a;
    for (int i = 1; i < argc; i++) {
        string k = argv[i];
        auto nxt = [&]() -> string { return i + 1 < argc ? argv[++i] : ""; };
        if (k == "--fcidump") a.fcidump = nxt();
        else if (k == "--pg") a.pg = nxt();
        else if (k == "--occ") a.occ = nxt();
        else if (k == "--bond") a.bond = atoi(nxt().c_str());
        else if (k == "--nsweeps") a.n_sweeps = atoi(nxt().c_str());
        else if (k == "--threads") a.threads = atoi(nxt().c_str());
        else if (k == "--seed") a.seed = atoi(nxt().c_str());
        else if (k == "--device") a.device = atoi(nxt().c_str());
        else if (k == "--scratch") a.scratch = nxt();
        else if (k == "--davidson") a.davidson = nxt();
        else if (k == "--conv") a.conv = atof(nxt().c_str());
        else if (k == "--noise") a.noise = atof(nxt().c_str());
        else if (k == "--dsize") a.dsize_gb = (size_t)atol(nxt().c_str());
        else if (k == "--compare") a.compare = true;
        else if (k == "--verify") a.verify = true;
        else if (k == "--gpu-rotate") a.gpu_rotate = true;   // default since round 2; kept for old command lines
        else if (k == "--gpu-contract") a.gpu_contract = true;
        else if (k == "--no-gpu-rotate") a.gpu_rotate = false;
        else if (k == "--no-gpu-contract") a.gpu_contract = false;
        else if (k == "--no-gpu-diag") a.gpu_diag = false;
        else if (k == "--no-gpu-iadd") a.gpu_iadd = false;
        else if (k == "--gpu-split") a.gpu_split = true; // density-matrix eigenproblems through b2g_syevd (cuSOLVER)
        else if (k == "--split-min-n") b2g_host::SplitHook::min_n() = atoi(nxt().c_str());
        else if (k == "--noise-sweeps") a.noise_sweeps = atoi(nxt().c_str());
        else if (k == "--dav-thrd") a.dav_thrd = atof(nxt().c_str());
        else if (k == "--restart-sweeps") a.restart_sweeps = atoi(nxt().c_str());
        else if (k == "--classic") a.classic = true; // ClassicParallelMPO instead of ParallelMPO (NewScheme)
        else if (k == "--cpu-only") a.cpu_only = true; // the stock CPU path alone (sweep-time baseline)
        else if (k == "--host-mirror") a.host_mirror = true;
        else if (k == "--no-pin") a.pin = false;
        else if (k == "--ranks") a.ranks = atoi(nxt().c_str());
        else if (k == "--rank") a.rank = atoi(nxt().c_str());
        else if (k == "--shm") a.shm = nxt();
        else {
            fprintf(stderr, "usage: b2g_dmrg --fcidump F [--pg d2h] [--bond M] [--nsweeps n] [--threads t] "
                            "[--davidson host|device] [--compare] [--verify] [--gpu-rotate] [--gpu-contract] [--occ F] [--noise x] [--conv x]\n");
            return 2;
        }
    }
    typedef B2G_S S;
    setvbuf(stdout, nullptr, _IOLBF, 0);
    Random::rand_seed(a.seed);
    frame_<double>() = make_shared<DataFrame<double>>((size_t)1 << 28, a.dsize_gb << 30, a.scratch);
    frame_<double>()->use_main_stack = false;
    frame_<double>()->minimal_disk_usage = true;
    threading_() = make_shared<Threading>(ThreadingTypes::OperatorBatchedGEMM | ThreadingTypes::Global, a.threads,
                                          a.threads, 1);
    threading_()->seq_type = SeqTypes::Tasked;
    shared_ptr<FCIDUMP<double>> fcidump = make_shared<FCIDUMP<double>>();
    fcidump->read(a.fcidump);
    PGTypes pg = pg_of(a.pg);
    vector<uint8_t> orbsym = fcidump->template orb_sym<uint8_t>();
    transform(orbsym.begin(), orbsym.end(), orbsym.begin(),
              [pg](uint8_t x) { return (uint8_t)PointGroup::swap_pg(pg)(x); });
    S vacuum(0);
    S target(fcidump->n_elec(), fcidump->twos(), PointGroup::swap_pg(pg)(fcidump->isym()));
    shared_ptr<HamiltonianQC<S, double>> hamil =
        make_shared<HamiltonianQC<S, double>>(vacuum, fcidump->n_sites(), orbsym, fcidump);
    shared_ptr<MPO<S, double>> mpo =
        make_shared<MPOQC<S, double>>(hamil, QCTypes::Conventional, "HQC", hamil->n_sites / 2 / 2 * 2);
    mpo->basis = hamil->basis;
    mpo = make_shared<SimplifiedMPO<S, double>>(mpo, make_shared<RuleQC<S, double>>(), true, true,
                                                OpNamesSet({OpNames::R, OpNames::RD}));
    // one process per GPU: the reference's own ParallelMPO over ParallelRuleQC (parallel_mpo.hpp:150,
    // qc_parallel_rule.hpp:44); host collectives through shared memory, sigma all-reduce over NCCL
    if (a.ranks > 1) {
        shared_ptr<ParallelCommunicator<S>> comm =
            make_shared<b2g_host::ShmCommunicator<S>>(a.ranks, a.rank, a.shm);
        shared_ptr<ParallelRule<S, double>> rule = make_shared<ParallelRuleQC<S, double>>(comm);
        // NewScheme (default, parallel_mpo.hpp:150): no blocking collectives, Partial operators repeated on every
        // rank.  --classic (parallel_mpo.hpp:32): every term on exactly one rank, Partial operators reduced to
        // their owner after blocking (ParallelRule::distributed_apply, parallel_rule.hpp:418-494).
        if (a.classic)
            mpo = make_shared<ClassicParallelMPO<S, double>>(mpo, rule);
        else
            mpo = make_shared<ParallelMPO<S, double>>(mpo, rule);
        // all ranks share the scratch directory: ParallelRule's constructor gives every rank its own
        // prefix for distributed files and lets only the root write the common ones (parallel_rule.hpp:340)
        if (a.rank != 0)
            cout.setstate(ios::failbit); // like MPICommunicator (parallel_mpi.hpp:60-61)
    }
    RunResult ref;
    vector<double> cont_ref, cont_gpu;
    if (a.compare) {
        printf("=== reference CPU path (stock TensorFunctions, %d threads) ===\n", a.threads);
        ref = run_dmrg<S>(a, mpo, hamil, target, false);
        if (a.restart_sweeps > 0) {
            printf("=== reference CPU path, %d zero-noise sweeps from its own final state ===\n", a.restart_sweeps);
            cont_ref = continue_dmrg<S>(a, mpo, false);
        }
    }
    if (a.cpu_only) {
        if (!a.compare)
            ref = run_dmrg<S>(a, mpo, hamil, target, false);
        for (size_t i = 0; i < ref.energies.size(); i++)
            printf("SWEEP %zu E_ref=%.12f\n", i, ref.energies[i]);
        printf("{\"mode\": \"b2g_dmrg\", \"cpu_only\": 1, \"bond\": %d, \"seed\": %d, \"sweeps\": %zu, \"t_ref\": %.3f, "
               "\"threads\": %d, \"e_ref\": %.12f}\n",
               a.bond, a.seed, ref.energies.size(), ref.total, a.threads, ref.energies.empty() ? 0.0 : ref.energies.back());
        fflush(stdout);
        _exit(0);
    }
    printf("=== GPU path (b2g_host::install, davidson = %s) ===\n", a.davidson.c_str());
    shared_ptr<TensorFunctions<S, double>> stock_tf = mpo->tf;
    shared_ptr<b2g_host::Session> session =
        a.ranks > 1 ? b2g_host::install_parallel<S>(mpo, a.device) : b2g_host::install<S>(mpo, a.device);
    session->verify = a.verify;
    session->gpu_rotate = a.gpu_rotate;
    session->gpu_contract = a.gpu_contract;
    session->gpu_diag = a.gpu_diag;
    session->gpu_iadd = a.gpu_iadd;
    session->host_mirror = a.host_mirror;
    session->arm_split(a.gpu_split);
    if (a.pin)
        session->pin_stacks();
    RunResult gpu = run_dmrg<S>(a, mpo, hamil, target, true);
    double restart_diff = 0;
    if (a.compare && a.restart_sweeps > 0) {
        printf("=== GPU path, %d zero-noise sweeps from the CPU arm's final state ===\n", a.restart_sweeps);
        cont_gpu = continue_dmrg<S>(a, mpo, true);
        for (size_t i = 0; i < min(cont_gpu.size(), cont_ref.size()); i++) {
            printf("RESTART SWEEP %zu E_gpu=%.12f E_ref=%.12f diff=%.3e\n", i, cont_gpu[i], cont_ref[i],
                   cont_gpu[i] - cont_ref[i]);
            restart_diff = max(restart_diff, fabs(cont_gpu[i] - cont_ref[i]));
        }
    }
    for (size_t i = 0; i < gpu.energies.size(); i++) {
        if (a.compare && i < ref.energies.size())
            printf("SWEEP %zu E_gpu=%.12f E_ref=%.12f diff=%.3e\n", i, gpu.energies[i], ref.energies[i],
                   gpu.energies[i] - ref.energies[i]);
        else
            printf("SWEEP %zu E_gpu=%.12f\n", i, gpu.energies[i]);
    }
    double maxdiff = 0;
    if (a.compare)
        for (size_t i = 0; i < min(gpu.energies.size(), ref.energies.size()); i++)
            maxdiff = max(maxdiff, fabs(gpu.energies[i] - ref.energies[i]));
    if (a.rank != 0) {
        fflush(stdout);
        _exit(0);
    }
    int64_t res_hit = 0, res_mirrored = 0;
    b2g_resident_stats(session->ctx, &res_hit, &res_mirrored);
    printf("{\"mode\": \"b2g_dmrg\", \"ranks\": %d, \"davidson\": \"%s\", \"bond\": %d, \"seed\": %d, \"sweeps\": %zu, "
           "\"t_gpu\": %.3f, \"t_ref\": %.3f, \"threads\": %d, \"e_gpu\": %.12f, \"e_ref\": %.12f, "
           "\"max_sweep_diff\": %.3e, \"final_diff\": %.3e, \"restart_sweeps\": %zu, \"max_restart_sweep_diff\": %.3e, "
           "\"plans\": %zu, \"host_matvecs\": %zu, \"t_plan\": %.3f, \"t_host_matvec\": %.3f, \"launches\": %lld, "
           "\"matvec_sites_verified\": %zu, \"max_matvec_rel_err\": %.3e, \"gpu_rotate\": %d, \"rotations\": %zu, "
           "\"t_rotate\": %.3f, \"t_rotate_download\": %.3f, \"rotate_gflop\": %.3f, \"max_rotate_rel_err\": %.3e, "
           "\"gpu_contract\": %d, \"contractions\": %zu, \"t_contract\": %.3f, \"contract_entries\": %zu, "
           "\"contract_kernel_ms\": %.3f, \"contract_gbytes\": %.4f, \"max_contract_rel_err\": %.3e, "
           "\"t_contract_record\": %.3f, \"t_contract_plan\": %.3f, \"t_contract_upload\": %.3f, "
           "\"t_contract_download\": %.3f, \"gpu_diag\": %d, \"diagonals\": %zu, \"t_diag\": %.3f, "
           "\"diag_entries\": %zu, \"max_diag_rel_err\": %.3e, \"gpu_iadd\": %d, \"iadd_walks\": %zu, \"t_iadd\": %.3f, "
           "\"iadd_entries\": %zu, \"max_iadd_rel_err\": %.3e, \"host_mirror\": %d, \"pinned_stacks\": %d, "
           "\"resident_read_gbytes\": %.3f, \"mirrored_gbytes\": %.3f, \"resident_peak_gbytes\": %.3f, "
           "\"resident_uploaded_gbytes\": %.3f, \"resident_downloaded_gbytes\": %.3f, \"resident_evicted_gbytes\": %.3f, "
           "\"t_precompute\": %.3f, \"t_davidson\": %.3f, \"t_contract_alloc\": %.3f, \"t_contract_ensure\": %.3f, "
           "\"t_contract_exec\": %.3f, \"t_rotate_alloc\": %.3f, \"t_rotate_exec\": %.3f, \"oom_retries\": %zu, \"gpu_split\": %d, \"syevd_gpu\": %zu, \"syevd_cpu\": %zu}\n",
           a.ranks, a.davidson.c_str(), a.bond, a.seed, gpu.energies.size(), gpu.total, ref.total, a.threads,
           gpu.energies.empty() ? 0.0 : gpu.energies.back(), ref.energies.empty() ? 0.0 : ref.energies.back(),
           maxdiff,
           (a.compare && !gpu.energies.empty() && !ref.energies.empty()) ? gpu.energies.back() - ref.energies.back() : 0.0,
           min(cont_gpu.size(), cont_ref.size()), restart_diff,
           session->n_plan, session->n_matvec, session->t_plan, session->t_matvec,
           (long long)b2g_context_launches(session->ctx), session->n_verified, session->max_matvec_err,
           (int)a.gpu_rotate, session->n_rotate, session->t_rotate, session->t_rotate_download,
           session->rotate_flops * 1e-9, session->max_rotate_err, (int)a.gpu_contract, session->n_contract,
           session->t_contract, session->contract_entries, session->contract_kernel_ms, session->contract_bytes * 1e-9,
           session->max_contract_err, session->t_contract_record, session->t_contract_plan,
           session->t_contract_upload, session->t_contract_download, (int)a.gpu_diag, session->n_diag,
           session->t_diag, session->diag_entries, session->max_diag_err, (int)a.gpu_iadd, session->n_iadd,
           session->t_iadd, session->iadd_entries, session->max_iadd_err, (int)session->host_mirror,
           (int)(session->pinned != nullptr), res_hit * 1e-9, res_mirrored * 1e-9, session->store->peak * 1e-9,
           session->store->uploaded_bytes * 1e-9, session->store->downloaded_bytes * 1e-9,
           session->store->evicted_bytes * 1e-9, session->t_precompute, session->t_davidson, session->t_contract_alloc,
           session->t_contract_ensure, session->t_contract_exec, session->t_rotate_alloc, session->t_rotate_exec,
           session->n_oom_retries, (int)session->gpu_split, b2g_host::SplitHook::calls_gpu().load(),
           b2g_host::SplitHook::calls_cpu().load());
    fflush(stdout);
    b2g_prof_dump(getenv("B2G_PROF_FILE")); // B2G_PROF: wall-clock sections of the library and the binding
    _exit(0);
}
