// b2g_shm_comm.hpp — single-node stand-in for block2's MPICommunicator (core/parallel_mpi.hpp:81)
// for one-process-per-GPU runs where no MPI is installed: the host-side collectives of the
// reference's parallel DMRG (Davidson control broadcasts, density-matrix / operator reductions,
// label all-gathers) go through a POSIX shared-memory segment; the sigma all-reduce of the H.C
// hot path itself runs on the GPUs over NCCL (b2g_allreduce_sum, inside b2g_seq_matvec /
// b2g_davidson).  Only the overloads the real-double ParallelRuleQC path calls are implemented;
// anything else keeps the base-class assert(size == 1).
#pragma once
#include "block2_core.hpp"
#include <atomic>
#include <fcntl.h>
#include <pthread.h>
#include <stdexcept>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

namespace b2g_host {

using namespace block2;

template <typename S> struct ShmCommunicator : ParallelCommunicator<S> {
    using ParallelCommunicator<S>::size;
    using ParallelCommunicator<S>::rank;
    using ParallelCommunicator<S>::root;
    struct Header {
        pthread_barrier_t bar;
        std::atomic<int> ready;
    };
    static constexpr size_t SLOT = (size_t)32 << 20;
    std::string name;
    Header *hdr = nullptr;
    char *slots = nullptr;
    size_t total = 0;
    ShmCommunicator(int size_, int rank_, const std::string &name_)
        : ParallelCommunicator<S>(size_, rank_, 0), name("/b2g_" + name_) {
        this->para_type = ParallelTypes::Distributed;
        total = sizeof(Header) + 4096 + SLOT * (size_t)size;
        int fd = -1;
        if (rank == 0) {
            shm_unlink(name.c_str());
            fd = shm_open(name.c_str(), O_CREAT | O_EXCL | O_RDWR, 0600);
            if (fd < 0 || ftruncate(fd, (off_t)total) != 0)
                throw std::runtime_error("ShmCommunicator: cannot create " + name);
        } else {
            for (int tries = 0; tries < 3000 && fd < 0; tries++) {
                fd = shm_open(name.c_str(), O_RDWR, 0600);
                if (fd < 0)
                    usleep(10000);
            }
            if (fd < 0)
                throw std::runtime_error("ShmCommunicator: cannot open " + name);
            struct stat st;
            while (fstat(fd, &st) == 0 && (size_t)st.st_size < total)
                usleep(1000);
        }
        void *p = mmap(nullptr, total, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
        close(fd);
        if (p == MAP_FAILED)
            throw std::runtime_error("ShmCommunicator: mmap failed");
        hdr = (Header *)p;
        slots = (char *)p + ((sizeof(Header) + 4095) & ~(size_t)4095);
        if (rank == 0) {
            pthread_barrierattr_t at;
            pthread_barrierattr_init(&at);
            pthread_barrierattr_setpshared(&at, PTHREAD_PROCESS_SHARED);
            pthread_barrier_init(&hdr->bar, &at, (unsigned)size);
            hdr->ready.store(1);
        } else
            while (hdr->ready.load() != 1)
                usleep(1000);
        barrier();
    }
    ~ShmCommunicator() override {
        if (hdr) {
            munmap((void *)hdr, total);
            if (rank == 0)
                shm_unlink(name.c_str());
        }
    }
    bool is_root() const noexcept override { return rank == root; }
    void barrier() override { pthread_barrier_wait(&hdr->bar); }
    char *slot(int r) const { return slots + SLOT * (size_t)r; }

    // element-wise reduction of typed arrays; owner < 0: every rank gets the result
    template <typename T, typename Op> void reduce(T *data, size_t len, int owner, Op op) {
        const size_t cap = SLOT / sizeof(T);
        for (size_t off = 0; off < len; off += cap) {
            const size_t n = std::min(cap, len - off);
            memcpy(slot(rank), data + off, n * sizeof(T));
            barrier();
            if (owner < 0 || owner == rank) {
                const T *s0 = (const T *)slot(0);
                for (size_t j = 0; j < n; j++)
                    data[off + j] = s0[j];
                for (int r = 1; r < size; r++) { // rank order: every rank computes identical bits
                    const T *sr = (const T *)slot(r);
                    for (size_t j = 0; j < n; j++)
                        data[off + j] = op(data[off + j], sr[j]);
                }
            }
            barrier();
        }
    }
    void bcast_bytes(void *data, size_t bytes, int owner) {
        for (size_t off = 0; off < bytes; off += SLOT) {
            const size_t n = std::min(SLOT, bytes - off);
            if (rank == owner)
                memcpy(slot(owner), (char *)data + off, n);
            barrier();
            if (rank != owner)
                memcpy((char *)data + off, slot(owner), n);
            barrier();
        }
    }
    struct Sum { template <typename T> T operator()(T a, T b) const { return a + b; } };
    struct Min { template <typename T> T operator()(T a, T b) const { return b < a ? b : a; } };
    struct Max { template <typename T> T operator()(T a, T b) const { return a < b ? b : a; } };
    struct Or  { char operator()(char a, char b) const { return (char)(a || b); } };
    struct Xor { char operator()(char a, char b) const { return (char)(a ^ b); } };

    void broadcast(double *d, size_t n, int o) override { bcast_bytes(d, n * sizeof(double), o); }
    void broadcast(long double *d, size_t n, int o) override { bcast_bytes(d, n * sizeof(long double), o); }
    void broadcast(int *d, size_t n, int o) override { bcast_bytes(d, n * sizeof(int), o); }
    void broadcast(long long int *d, size_t n, int o) override { bcast_bytes(d, n * sizeof(long long int), o); }
    void broadcast(const shared_ptr<SparseMatrix<S, double>> &m, int o) override {
        bcast_bytes(m->data, m->total_memory * sizeof(double), o);
    }
    void ibroadcast(const shared_ptr<SparseMatrix<S, double>> &m, int o) override { broadcast(m, o); }
    void ibroadcast(double *d, size_t n, int o) override { broadcast(d, n, o); }
    void allreduce_sum(double *d, size_t n) override { reduce(d, n, -1, Sum()); }
    void allreduce_sum(const shared_ptr<SparseMatrix<S, double>> &m) override {
        reduce(m->data, m->total_memory, -1, Sum());
    }
    void allreduce_sum(const shared_ptr<SparseMatrixGroup<S, double>> &m) override {
        reduce(m->data, m->total_memory, -1, Sum());
    }
    void allreduce_sum(vector<S> &vs) override { // all-gather of quantum labels (parallel_mpi.hpp:541-557)
        long long int n = (long long int)vs.size(), nmax = n;
        reduce(&nmax, 1, -1, Max());
        vector<S> mine(vs);
        mine.resize((size_t)nmax, S(S::invalid));
        if ((size_t)nmax * sizeof(S) > SLOT)
            throw std::runtime_error("ShmCommunicator: label list too long");
        memcpy(slot(rank), mine.data(), (size_t)nmax * sizeof(S));
        barrier();
        vector<S> all;
        for (int r = 0; r < size; r++) {
            const S *p = (const S *)slot(r);
            for (long long int j = 0; j < nmax; j++)
                if (!(p[j] == S(S::invalid)))
                    all.push_back(p[j]);
        }
        barrier();
        vs = all;
    }
    void allreduce_min(double *d, size_t n) override { reduce(d, n, -1, Min()); }
    void allreduce_min(long double *d, size_t n) override { reduce(d, n, -1, Min()); }
    void allreduce_min(vector<double> &v) override { reduce(v.data(), v.size(), -1, Min()); }
    void allreduce_min(vector<long double> &v) override { reduce(v.data(), v.size(), -1, Min()); }
    void allreduce_min(vector<vector<double>> &vs) override {
        for (auto &v : vs)
            reduce(v.data(), v.size(), -1, Min());
    }
    void allreduce_min(vector<vector<long double>> &vs) override {
        for (auto &v : vs)
            reduce(v.data(), v.size(), -1, Min());
    }
    void allreduce_max(double *d, size_t n) override { reduce(d, n, -1, Max()); }
    void allreduce_max(vector<double> &v) override { reduce(v.data(), v.size(), -1, Max()); }
    void allreduce_logical_or(char *d, size_t n) override { reduce(d, n, -1, Or()); }
    void allreduce_xor(char *d, size_t n) override { reduce(d, n, -1, Xor()); }
    void reduce_sum(double *d, size_t n, int o) override { reduce(d, n, o, Sum()); }
    void reduce_sum(uint64_t *d, size_t n, int o) override { reduce(d, n, o, Sum()); }
    void reduce_sum_optional(double *d, size_t n, int o) override { reduce(d, n, o, Sum()); }
    void reduce_sum_optional(uint64_t *d, size_t n, int o) override { reduce(d, n, o, Sum()); }
    void ireduce_sum(double *d, size_t n, int o) override { reduce(d, n, o, Sum()); }
    void reduce_sum(const shared_ptr<SparseMatrix<S, double>> &m, int o) override {
        reduce(m->data, m->total_memory, o, Sum());
    }
    void ireduce_sum(const shared_ptr<SparseMatrix<S, double>> &m, int o) override { reduce_sum(m, o); }
    void reduce_sum(const shared_ptr<SparseMatrixGroup<S, double>> &m, int o) override {
        reduce(m->data, m->total_memory, o, Sum());
    }
    void reduce_max(uint64_t *d, size_t n, int o) override { reduce(d, n, o, Max()); }
    void waitall() override {}
};

} // namespace b2g_host
