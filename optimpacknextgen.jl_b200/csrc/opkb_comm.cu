// opkb_comm.cu -- the only communication of the hot path: one small exchange of the
// per-rank reduction partials per phase: a shared-memory mailbox the last CTA
// of every rank writes into (ranks of one node), or a packed all-gather over NCCL
// (NVLink 5 / NVSwitch); plus the point-to-point halo rows of the stencil objective.  NCCL is loaded with dlopen so that the single-GPU
// library has no link-time dependency; inside a torch process this resolves to
// the libnccl.so.2 torch has already loaded.
#include <dlfcn.h>
#include <fcntl.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <time.h>
#include <unistd.h>

#include "opkb_internal.cuh"

// minimal NCCL ABI (nccl.h, stable since 2.x)
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
enum { ncclFloat32 = 7, ncclFloat64 = 8 };

static struct {
    void* handle;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*);
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int);
    ncclResult_t (*CommDestroy)(ncclComm_t);
    ncclResult_t (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t);
    const char* (*GetErrorString)(ncclResult_t);
    ncclResult_t (*Send)(const void*, size_t, int, int, ncclComm_t, cudaStream_t);
    ncclResult_t (*Recv)(void*, size_t, int, int, ncclComm_t, cudaStream_t);
    ncclResult_t (*GroupStart)(void);
    ncclResult_t (*GroupEnd)(void);
} g_nccl;

static int load_nccl(void)
{
    if (g_nccl.handle) return 0;
    const char* names[] = {"libnccl.so.2", "libnccl.so", NULL};
    void* h = NULL;
    for (int i = 0; names[i] && !h; ++i) h = dlopen(names[i], RTLD_NOW | RTLD_GLOBAL);
    if (!h) return opkb_set_error(OPKB_ENCCL, "cannot load libnccl.so.2: %s", dlerror());
    g_nccl.GetUniqueId = (ncclResult_t(*)(ncclUniqueId*))dlsym(h, "ncclGetUniqueId");
    g_nccl.CommInitRank = (ncclResult_t(*)(ncclComm_t*, int, ncclUniqueId, int))dlsym(h, "ncclCommInitRank");
    g_nccl.CommDestroy = (ncclResult_t(*)(ncclComm_t))dlsym(h, "ncclCommDestroy");
    g_nccl.AllGather = (ncclResult_t(*)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t))
        dlsym(h, "ncclAllGather");
    g_nccl.GetErrorString = (const char* (*)(ncclResult_t))dlsym(h, "ncclGetErrorString");
    g_nccl.Send = (ncclResult_t(*)(const void*, size_t, int, int, ncclComm_t, cudaStream_t))dlsym(h, "ncclSend");
    g_nccl.Recv = (ncclResult_t(*)(void*, size_t, int, int, ncclComm_t, cudaStream_t))dlsym(h, "ncclRecv");
    g_nccl.GroupStart = (ncclResult_t(*)(void))dlsym(h, "ncclGroupStart");
    g_nccl.GroupEnd = (ncclResult_t(*)(void))dlsym(h, "ncclGroupEnd");
    if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.AllGather || !g_nccl.GetErrorString)
        return opkb_set_error(OPKB_ENCCL, "libnccl lacks a required symbol");
    g_nccl.handle = h;
    return 0;
}

static int nccl_fail(ncclResult_t r, const char* what)
{
    return opkb_set_error(OPKB_ENCCL, "NCCL error %d (%s) in %s", (int)r,
                          g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?", what);
}

extern "C" int opkb_comm_unique_id(void* id128)
{
    if (!id128) return opkb_set_error(OPKB_EINVAL, "NULL id buffer");
    int rc = load_nccl();
    if (rc) return rc;
    ncclUniqueId id;
    ncclResult_t r = g_nccl.GetUniqueId(&id);
    if (r != 0) return nccl_fail(r, "ncclGetUniqueId");
    memcpy(id128, &id, sizeof(id));
    return 0;
}

// Shared-memory mailbox of the ranks of one node (see opkb_context_).  The segment is named
// after the NCCL unique id (the one thing every rank already shares), created by rank 0, mapped
// and registered with CUDA by every rank, and unlinked once all ranks hold it.  Returns 1 if
// this rank holds a registered mapping, 0 otherwise (the caller makes the ranks agree).
static int mailbox_open(opkb_context* ctx, const void* id128, int rank, int nranks, void** host, void** dev,
                        size_t* bytes_out, char* name_out)
{
    const unsigned char* id = (const unsigned char*)id128;
    unsigned long long h = 1469598103934665603ull;
    for (int i = 0; i < 128; ++i) h = (h ^ id[i]) * 1099511628211ull;
    snprintf(name_out, 64, "/opkb200_%016llx", h);
    const size_t bytes = (size_t)nranks * OPKB_MBOX_SLOT;
    int fd = -1;
    if (rank == 0) {
        shm_unlink(name_out);
        fd = shm_open(name_out, O_CREAT | O_EXCL | O_RDWR, 0600);
        if (fd >= 0 && ftruncate(fd, (off_t)bytes) != 0) {
            close(fd);
            fd = -1;
        }
    } else {
        for (int tries = 0; tries < 30000 && fd < 0; ++tries) {        // up to ~30 s
            fd = shm_open(name_out, O_RDWR, 0600);
            struct stat st;
            if (fd >= 0 && (fstat(fd, &st) != 0 || (size_t)st.st_size < bytes)) {
                close(fd);
                fd = -1;
            }
            if (fd < 0) {
                struct timespec ts = {0, 1000000};
                nanosleep(&ts, NULL);
            }
        }
    }
    if (fd < 0) return 0;
    void* p = mmap(NULL, bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    close(fd);
    if (p == MAP_FAILED) return 0;
    if (cudaHostRegister(p, bytes, cudaHostRegisterMapped | cudaHostRegisterPortable) != cudaSuccess) {
        cudaGetLastError();
        munmap(p, bytes);
        return 0;
    }
    void* d = NULL;
    if (cudaHostGetDevicePointer(&d, p, 0) != cudaSuccess) {
        cudaGetLastError();
        cudaHostUnregister(p);
        munmap(p, bytes);
        return 0;
    }
    (void)ctx;
    *host = p;
    *dev = d;
    *bytes_out = bytes;
    return 1;
}

extern "C" int opkb_comm_init(opkb_context* ctx, const void* id128, int rank, int nranks)
{
    if (!ctx || !id128) return opkb_set_error(OPKB_EINVAL, "NULL argument");
    if (nranks < 1 || rank < 0 || rank >= nranks) return opkb_set_error(OPKB_EINVAL, "invalid rank");
    if (nranks == 1) {
        ctx->rank = 0;
        ctx->nranks = 1;
        return 0;
    }
    int rc = load_nccl();
    if (rc) return rc;
    OPKB_CUDA(cudaSetDevice(ctx->device));
    ncclUniqueId id;
    memcpy(&id, id128, sizeof(id));
    ncclComm_t comm = NULL;
    ncclResult_t r = g_nccl.CommInitRank(&comm, nranks, id, rank);
    if (r != 0) return nccl_fail(r, "ncclCommInitRank");
    ctx->nccl_comm = comm;
    ctx->rank = rank;
    ctx->nranks = nranks;
    OPKB_CUDA(cudaMalloc(&ctx->d_all, sizeof(double) * OPKB_NRED_MAX * (size_t)nranks));
    OPKB_CUDA(cudaStreamSynchronize(ctx->stream));
    ctx->d_res = ctx->d_res_dev;   // NCCL sends from device memory
    cudaFreeHost(ctx->h_res);
    OPKB_CUDA(cudaMallocHost(&ctx->h_res, sizeof(double) * OPKB_NRED_MAX * (size_t)nranks));
    // The per-phase exchange of the reduced scalars: a shared-memory mailbox when every rank of
    // the communicator can map it (one node), the packed ncclAllGather otherwise / OPKB_MAILBOX=0.
    const char* e = getenv("OPKB_MAILBOX");
    void *host = NULL, *dev = NULL;
    size_t bytes = 0;
    char name[64] = "";
    int mine = 0;
    if (!e || atoi(e) != 0) mine = mailbox_open(ctx, id128, rank, nranks, &host, &dev, &bytes, name);
    // agreement (and the barrier after which the name can go): all-gather of the flags
    double flag = (double)mine;
    OPKB_CUDA(cudaMemcpyAsync(ctx->d_res_dev, &flag, sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    int rc2 = opkb_nccl_allgather_f64(ctx, ctx->d_res_dev, ctx->d_all, 1);
    if (rc2) return rc2;
    OPKB_CUDA(cudaMemcpyAsync(ctx->h_res, ctx->d_all, sizeof(double) * nranks, cudaMemcpyDeviceToHost, ctx->stream));
    OPKB_CUDA(cudaStreamSynchronize(ctx->stream));
    int all = 1;
    for (int q = 0; q < nranks; ++q) all = all && (ctx->h_res[q] == 1.0);
    if (rank == 0 && name[0]) shm_unlink(name);
    if (all) {
        ctx->mbox = (unsigned char*)host;
        ctx->mbox_dev = (unsigned char*)dev;
        ctx->mbox_bytes = bytes;
        ctx->phase = 0;
    } else if (mine) {
        cudaHostUnregister(host);
        munmap(host, bytes);
    }
    return 0;
}

extern "C" int opkb_comm_uses_mailbox(const opkb_context* ctx) { return (ctx && ctx->mbox) ? 1 : 0; }

void opkb_mailbox_close(opkb_context* ctx)
{
    if (!ctx || !ctx->mbox) return;
    cudaHostUnregister(ctx->mbox);
    munmap(ctx->mbox, ctx->mbox_bytes);
    ctx->mbox = ctx->mbox_dev = NULL;
}

extern "C" int opkb_comm_rank(const opkb_context* ctx) { return ctx ? ctx->rank : -1; }
extern "C" int opkb_comm_size(const opkb_context* ctx) { return ctx ? ctx->nranks : -1; }

int opkb_nccl_allgather_f64(opkb_context* ctx, const double* send, double* recv, int count)
{
    if (!ctx->nccl_comm) return opkb_set_error(OPKB_ENCCL, "communicator not initialised");
    ncclResult_t r = g_nccl.AllGather(send, recv, (size_t)count, ncclFloat64,
                                      (ncclComm_t)ctx->nccl_comm, ctx->stream);
    if (r != 0) return nccl_fail(r, "ncclAllGather");
    return 0;
}

// The one real data exchange of the library: the boundary rows of a
// row-sharded grid go to the two neighbouring ranks, point to point over NVLink, as
// ONE NCCL group (both directions at once) on the given stream.
int opkb_nccl_halo(opkb_context* ctx, int eltype, size_t count, const void* send_prev, void* recv_prev,
                   const void* send_next, void* recv_next, cudaStream_t stream)
{
    if (!ctx->nccl_comm) return opkb_set_error(OPKB_ENCCL, "communicator not initialised");
    if (!g_nccl.Send || !g_nccl.Recv || !g_nccl.GroupStart || !g_nccl.GroupEnd)
        return opkb_set_error(OPKB_ENCCL, "libnccl lacks ncclSend/ncclRecv");
    const int dt = (eltype == OPKB_DOUBLE) ? ncclFloat64 : ncclFloat32;
    ncclComm_t comm = (ncclComm_t)ctx->nccl_comm;
    ncclResult_t r = g_nccl.GroupStart();
    if (r != 0) return nccl_fail(r, "ncclGroupStart");
    if (ctx->rank > 0) {
        if (r == 0 && send_prev) r = g_nccl.Send(send_prev, count, dt, ctx->rank - 1, comm, stream);
        if (r == 0 && recv_prev) r = g_nccl.Recv(recv_prev, count, dt, ctx->rank - 1, comm, stream);
    }
    if (ctx->rank < ctx->nranks - 1) {
        if (r == 0 && send_next) r = g_nccl.Send(send_next, count, dt, ctx->rank + 1, comm, stream);
        if (r == 0 && recv_next) r = g_nccl.Recv(recv_next, count, dt, ctx->rank + 1, comm, stream);
    }
    ncclResult_t r2 = g_nccl.GroupEnd();
    if (r != 0) return nccl_fail(r, "ncclSend/ncclRecv");
    if (r2 != 0) return nccl_fail(r2, "ncclGroupEnd");
    return 0;
}
