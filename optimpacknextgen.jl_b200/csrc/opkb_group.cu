// opkb_group.cu -- ONE process, G devices ("One process, G devices, one stream per
// device"): a group owns one context per GPU, the contexts are the ranks 0..G-1 of an in-process
// communicator, and the native solvers of all ranks run concurrently on one host thread per GPU
// INSIDE the library.  A single `vmlmb(fg!, x0; lower, upper)` call of a single host process
// (src/quasinewton.jl:215, :226-242) then uses the whole box; the host shards its vectors in
// contiguous blocks over the contexts exactly as the ranks of a multi-process job do, the same
// kernels run, and the packed partials of every phase meet in the same mailbox layout -- here a
// pinned, portable, device-mapped host segment of the process instead of a POSIX shared-memory
// segment -- combined in rank order: bit-identical to the G-process run.
#include <stdlib.h>
#include <string.h>

#include <string>
#include <thread>
#include <vector>

#include "opkb_internal.cuh"

struct opkb_group_ {
    int n;
    opkb_context** ctx;
    unsigned char* mbox;     // cudaHostAlloc(portable | mapped): one segment for all ranks
    size_t mbox_bytes;
};

extern "C" int opkb_group_destroy(opkb_group* g)
{
    if (!g) return 0;
    for (int r = 0; r < g->n; ++r) {
        if (!g->ctx || !g->ctx[r]) continue;
        // the segment belongs to the group: the contexts must not unmap it
        g->ctx[r]->mbox = g->ctx[r]->mbox_dev = NULL;
        opkb_context_destroy(g->ctx[r]);
    }
    if (g->mbox) cudaFreeHost(g->mbox);
    free(g->ctx);
    free(g);
    return 0;
}

extern "C" int opkb_group_create(opkb_group** out, int ndev, const int* devices)
{
    if (!out) return opkb_set_error(OPKB_EINVAL, "NULL group pointer");
    *out = NULL;
    int have = 0;
    cudaError_t err = cudaGetDeviceCount(&have);
    if (err != cudaSuccess || have <= 0)
        return opkb_set_error(OPKB_ECUDA, "no CUDA device available (%s): libopkb200 has no CPU fallback",
                              cudaGetErrorString(err));
    if (ndev <= 0) ndev = have;
    if (ndev > have && !devices) return opkb_set_error(OPKB_EINVAL, "%d devices asked for, %d present", ndev, have);
    opkb_group* g = (opkb_group*)calloc(1, sizeof(opkb_group));
    if (!g) return opkb_set_error(OPKB_ENOMEM, "out of host memory");
    g->n = ndev;
    g->ctx = (opkb_context**)calloc((size_t)ndev, sizeof(opkb_context*));
    if (!g->ctx) { free(g); return opkb_set_error(OPKB_ENOMEM, "out of host memory"); }
    int rc = 0;
    for (int r = 0; r < ndev && !rc; ++r) rc = opkb_context_create(&g->ctx[r], devices ? devices[r] : r);
    if (!rc && ndev > 1) {
        g->mbox_bytes = (size_t)ndev * OPKB_MBOX_SLOT;
        err = cudaHostAlloc((void**)&g->mbox, g->mbox_bytes, cudaHostAllocPortable | cudaHostAllocMapped);
        if (err != cudaSuccess) rc = opkb_check_cuda(err, "cudaHostAlloc(mailbox)");
        else memset(g->mbox, 0, g->mbox_bytes);
    }
    for (int r = 0; r < ndev && !rc && ndev > 1; ++r) {
        opkb_context* c = g->ctx[r];
        cudaSetDevice(c->device);
        void* d = NULL;
        err = cudaHostGetDevicePointer(&d, g->mbox, 0);
        if (err != cudaSuccess) { rc = opkb_check_cuda(err, "cudaHostGetDevicePointer(mailbox)"); break; }
        c->rank = r;
        c->nranks = ndev;
        c->mbox = g->mbox;
        c->mbox_dev = (unsigned char*)d;
        c->mbox_bytes = g->mbox_bytes;
        c->phase = 0;
        c->d_res = c->d_res_dev;
    }
    if (rc) {
        opkb_group_destroy(g);
        return rc;
    }
    *out = g;
    return 0;
}

extern "C" int opkb_group_size(const opkb_group* g) { return g ? g->n : 0; }

extern "C" opkb_context* opkb_group_context(opkb_group* g, int rank)
{
    return (g && rank >= 0 && rank < g->n) ? g->ctx[rank] : NULL;
}

// fn(rank) on one host thread per GPU; the first failing rank's status and message are the caller's
template <typename F>
static int group_run(opkb_group* g, F fn)
{
    const int n = g->n;
    std::vector<int> rcs((size_t)n, 0);
    std::vector<std::string> msgs((size_t)n);
    auto body = [&](int r) {
        cudaSetDevice(g->ctx[r]->device);
        rcs[r] = fn(r);
        if (rcs[r]) msgs[r] = opkb_last_error();   // the message is thread-local
    };
    std::vector<std::thread> th;
    for (int r = 1; r < n; ++r) th.emplace_back(body, r);
    body(0);
    for (auto& t : th) t.join();
    for (int r = 0; r < n; ++r)
        if (rcs[r]) return opkb_set_error(rcs[r], "rank %d: %s", r, msgs[r].c_str());
    return 0;
}

extern "C" int opkb_group_solver_run(opkb_group* g, opkb_solver* const* solvers, int64_t niter, int* tasks)
{
    if (!g || !solvers || !tasks) return opkb_set_error(OPKB_EINVAL, "NULL argument");
    for (int r = 0; r < g->n; ++r)
        if (!solvers[r]) return opkb_set_error(OPKB_EINVAL, "NULL solver of rank %d", r);
    return group_run(g, [&](int r) { return opkb_solver_run(solvers[r], niter, &tasks[r]); });
}

extern "C" int opkb_group_spg_solver_run(opkb_group* g, opkb_spg_solver* const* solvers, int64_t niter,
                                         int* searching)
{
    if (!g || !solvers || !searching) return opkb_set_error(OPKB_EINVAL, "NULL argument");
    for (int r = 0; r < g->n; ++r)
        if (!solvers[r]) return opkb_set_error(OPKB_EINVAL, "NULL solver of rank %d", r);
    return group_run(g, [&](int r) { return opkb_spg_solver_run(solvers[r], niter, &searching[r]); });
}

// Any other call that returns reduced scalars blocks until every rank has made it; a host that
// wants to drive the ranks itself (Tier 1 / Tier 2 calls) runs `fn(rank, user)` on the group's
// threads: the ranks then advance in lock step like the processes of a torchrun job.
extern "C" int opkb_group_run(opkb_group* g, int (*fn)(int rank, void* user), void* user)
{
    if (!g || !fn) return opkb_set_error(OPKB_EINVAL, "NULL argument");
    return group_run(g, [&](int r) { return fn(r, user); });
}
