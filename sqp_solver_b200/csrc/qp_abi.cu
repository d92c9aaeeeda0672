// C-ABI implementation (host side) of include/sqp_b200_qp.h.
// Owns device state for a batch of solver instances, stages host buffers in pipelined chunks,
// picks a kernel, and launches it.  No exceptions cross this boundary.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "qp_common.cuh"
#include "qp_tile.cuh"

using namespace sqpb200;

static thread_local std::string g_create_error;

struct sqpb200_ctx {
    int device = 0;
    cudaDeviceProp prop{};
    cudaStream_t stream = nullptr;       // internal housekeeping stream (state initialisation)
    cudaStream_t copy_stream = nullptr;  // H2D staging for HOST_PTRS calls
    int *counters = nullptr;             // ring of work-queue counters, one per launch
    static constexpr int kCounters = 1024;
    long long launches = 0;
    int opt_kernel = 0, opt_chunks = 16, opt_ctas_per_sm = 0, opt_tile_warps = 0;
    int opt_slice = -1;  // time slicing of the register-tiled kernel: -1 = automatic, 0 = off, > 0 = iterations per slice
    std::string err;
    char last_kernel[64] = "none";
};

struct sqpb200_qp_batch {
    sqpb200_ctx *ctx = nullptr;
    int batch = 0, n = 0, m = 0;
    double *x = nullptr, *y = nullptr, *z = nullptr;
    int *status = nullptr, *iter = nullptr, *rho_updates = nullptr;
    double *rho_estimate = nullptr, *res_prim = nullptr, *res_dual = nullptr, *rho = nullptr;
    signed char *ctype = nullptr;
    double *fact = nullptr;  // lazily allocated
    size_t fact_doubles = 0;  // per instance
    double *fact_rho = nullptr;
    bool fused_used = false;
    int f32 = 0;  // compute precision of the register-tiled kernel for this batch (QPSolver<float>)
    bool fact_valid = false;  // a setup()/update_qp()/solve() launch has stored H^-1, rho and classes
    int fact_kernel = 0;      // which kernel family wrote the stored factor (its layout differs per kernel): KERNEL_* below, 0 = none
    int keep_kernel = 0;      // which kernel family wrote the factor kept by SQPB200_KEEP_FACTOR
    // HOST_PTRS staging protocol, per batch object (two objects of one context may stage concurrently): the device flag holding
    // the number of QPs whose inputs have landed, the pinned chunk boundaries the flag copies read, and the ordering event
    int *ready_dev = nullptr, *ready_host = nullptr;
    cudaEvent_t stage_event = nullptr;
    cudaStream_t last_stream = nullptr;  // stream of the last launch on this object, and an event recorded behind it:
    cudaEvent_t last_event = nullptr;    // get / set_iterates / total_iters on ANOTHER stream wait for it first
    int *rq = nullptr;  // time slicing: re-queue ring + [rq_cap] = slots handed out, [rq_cap + 1] = QPs finished
    int rq_cap = 0;
    double *gen_scratch = nullptr;  // generic kernel: per-CTA n*n factorisation workspace (owned by the batch object: launches of
    size_t gen_scratch_bytes = 0;   // different batch objects may overlap on different streams)
    unsigned long long *total_iters = nullptr;
    // staging for HOST_PTRS calls (lazily allocated)
    double *dP = nullptr, *dq = nullptr, *dA = nullptr, *dl = nullptr, *du = nullptr;
    // sparse-A entry point: device copies of the shared pattern and of the per-instance values
    int *sp_outer = nullptr, *sp_inner = nullptr;
    double *sp_vals = nullptr;
    int sp_nnz_cap = 0;
    int *sp2_outer = nullptr, *sp2_inner = nullptr, *sp2_perm = nullptr;  // the other compressed view of the pattern
    unsigned long long sp_hash = 0;  // hash of the pattern whose derived views (sp2_*, sp_pack) are on the device; 0 = none
    double *cl_scratch = nullptr;  // cluster kernel: per-cluster exchange buffers (owned by the batch object: launches of different
    size_t cl_scratch_bytes = 0;   // batch objects may overlap on different streams)
    unsigned *sp_pack = nullptr;  // [2][cap]: packed (index | value position << PACK_BITS) entries of the CSC and the CSR view (cluster kernel)
};

static int fail(sqpb200_ctx *ctx, int code, const char *what, cudaError_t e = cudaSuccess) {
    char buf[512];
    if (e != cudaSuccess)
        snprintf(buf, sizeof buf, "%s: %s (%s)", what, cudaGetErrorName(e), cudaGetErrorString(e));
    else
        snprintf(buf, sizeof buf, "%s", what);
    if (ctx) ctx->err = buf;
    else g_create_error = buf;
    return code;
}
#define CK(ctx, call)                                                        \
    do {                                                                     \
        cudaError_t e__ = (call);                                            \
        if (e__ != cudaSuccess) return fail((ctx), SQPB200_ERR_CUDA, #call, e__); \
    } while (0)

extern "C" {

int sqpb200_abi_version(void) { return SQPB200_ABI_VERSION; }

void sqpb200_qp_default_settings(sqpb200_qp_settings *s) {  // qp.hpp:38-53
    if (!s) return;
    s->rho = 1e-1;
    s->sigma = 1e-6;
    s->alpha = 1.0;
    s->eps_rel = 1e-3;
    s->eps_abs = 1e-3;
    s->max_iter = 1000;
    s->check_termination = 25;
    s->warm_start = 0;
    s->adaptive_rho = 0;
    s->adaptive_rho_tolerance = 5;
    s->adaptive_rho_interval = 25;
    s->verbose = 0;
}

int sqpb200_constr_type_init(const double *l, const double *u, int m, int *constr_type) {  // qp.cpp:283-294
    if (!l || !u || !constr_type || m < 0) return SQPB200_ERR_INVALID;
    for (int i = 0; i < m; i++) {
        if (l[i] < -LOOSE_BOUNDS_THRESH && u[i] > LOOSE_BOUNDS_THRESH) constr_type[i] = SQPB200_LOOSE_BOUNDS;
        else if (u[i] - l[i] < RHO_TOL) constr_type[i] = SQPB200_EQUALITY_CONSTRAINT;
        else constr_type[i] = SQPB200_INEQUALITY_CONSTRAINT;
    }
    return SQPB200_OK;
}

int sqpb200_ctx_create(int device, sqpb200_ctx **out) {
    if (!out) return fail(nullptr, SQPB200_ERR_INVALID, "sqpb200_ctx_create: out is NULL");
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(nullptr, SQPB200_ERR_CUDA, "sqpb200_ctx_create: no CUDA device (this library has no CPU fallback)", e);
    if (device < 0 || device >= ndev) return fail(nullptr, SQPB200_ERR_INVALID, "sqpb200_ctx_create: bad device index");
    sqpb200_ctx *c = new (std::nothrow) sqpb200_ctx();
    if (!c) return fail(nullptr, SQPB200_ERR_NOMEM, "sqpb200_ctx_create: host allocation failed");
    c->device = device;
    if ((e = cudaSetDevice(device)) != cudaSuccess || (e = cudaGetDeviceProperties(&c->prop, device)) != cudaSuccess ||
        (e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking)) != cudaSuccess ||
        (e = cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking)) != cudaSuccess ||
        (e = cudaMalloc(&c->counters, sizeof(int) * sqpb200_ctx::kCounters)) != cudaSuccess) {
        int rc = fail(nullptr, SQPB200_ERR_CUDA, "sqpb200_ctx_create", e);
        delete c;
        return rc;
    }
    if (c->prop.major < 10) {
        delete c;
        return fail(nullptr, SQPB200_ERR_UNSUPPORTED, "sqpb200_ctx_create: kernels are built for sm_100a only");
    }
    *out = c;
    return SQPB200_OK;
}

int sqpb200_ctx_destroy(sqpb200_ctx *c) {
    if (!c) return SQPB200_OK;
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    if (c->counters) cudaFree(c->counters);
    if (c->stream) cudaStreamDestroy(c->stream);
    if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
    delete c;
    return SQPB200_OK;
}

int sqpb200_ctx_set_option(sqpb200_ctx *c, int option, int value) {
    if (!c) return SQPB200_ERR_INVALID;
    switch (option) {
        case SQPB200_OPT_KERNEL:
            if (value < 0 || value > 5) return fail(c, SQPB200_ERR_INVALID, "SQPB200_OPT_KERNEL: value must be 0..5");
            c->opt_kernel = value;
            return SQPB200_OK;
        case SQPB200_OPT_H2D_CHUNKS:
            if (value < 1 || value > 64) return fail(c, SQPB200_ERR_INVALID, "SQPB200_OPT_H2D_CHUNKS: 1..64");
            c->opt_chunks = value;
            return SQPB200_OK;
        case SQPB200_OPT_CTAS_PER_SM:
            if (value < 0 || value > 32) return fail(c, SQPB200_ERR_INVALID, "SQPB200_OPT_CTAS_PER_SM: 0..32");
            c->opt_ctas_per_sm = value;
            return SQPB200_OK;
        case SQPB200_OPT_SLICE_ITERS:
            if (value < -1) return fail(c, SQPB200_ERR_INVALID, "SQPB200_OPT_SLICE_ITERS: -1 (automatic), 0 (off) or iterations per slice (+ 65536 x iterations of the first slice)");
            c->opt_slice = value;
            return SQPB200_OK;
        case SQPB200_OPT_TILE_WARPS:
            if (value != 0 && value != 1 && value != 2 && value != 4 && value != 8)
                return fail(c, SQPB200_ERR_INVALID, "SQPB200_OPT_TILE_WARPS: 0, 1, 2, 4 or 8");
            c->opt_tile_warps = value;
            return SQPB200_OK;
    }
    return fail(c, SQPB200_ERR_INVALID, "sqpb200_ctx_set_option: unknown option");
}

const char *sqpb200_last_error(const sqpb200_ctx *c) { return c ? c->err.c_str() : g_create_error.c_str(); }

int sqpb200_device_query(const sqpb200_ctx *c, int *device, int *sm_count, int *cc_major, int *cc_minor,
                         size_t *smem_per_block_optin) {
    if (!c) return SQPB200_ERR_INVALID;
    if (device) *device = c->device;
    if (sm_count) *sm_count = c->prop.multiProcessorCount;
    if (cc_major) *cc_major = c->prop.major;
    if (cc_minor) *cc_minor = c->prop.minor;
    if (smem_per_block_optin) *smem_per_block_optin = c->prop.sharedMemPerBlockOptin;
    return SQPB200_OK;
}

int sqpb200_measure_fp64_peak(sqpb200_ctx *c, double *tflops, double *seconds, double *fma_count) {
    if (!c || !tflops) return SQPB200_ERR_INVALID;
    CK(c, cudaSetDevice(c->device));
    double sec = 0, cnt = 0;
    CK(c, measure_dfma_peak(c->prop.multiProcessorCount, c->stream, &sec, &cnt));
    c->launches += 5;
    *tflops = 2.0 * cnt / sec / 1e12;
    if (seconds) *seconds = sec;
    if (fma_count) *fma_count = cnt;
    return SQPB200_OK;
}

long long sqpb200_launch_count(const sqpb200_ctx *c) { return c ? c->launches : 0; }
const char *sqpb200_last_kernel(const sqpb200_ctx *c) { return c ? c->last_kernel : "none"; }

int sqpb200_dev_alloc(sqpb200_ctx *c, size_t bytes, void **dev_ptr) {
    if (!c || !dev_ptr) return SQPB200_ERR_INVALID;
    *dev_ptr = nullptr;
    CK(c, cudaSetDevice(c->device));
    cudaError_t e = cudaMalloc(dev_ptr, bytes ? bytes : 1);
    if (e != cudaSuccess) return fail(c, SQPB200_ERR_NOMEM, "sqpb200_dev_alloc", e);
    return SQPB200_OK;
}
int sqpb200_dev_free(sqpb200_ctx *c, void *dev_ptr) {
    if (!c) return SQPB200_ERR_INVALID;
    CK(c, cudaSetDevice(c->device));
    if (dev_ptr) CK(c, cudaFree(dev_ptr));
    return SQPB200_OK;
}
int sqpb200_dev_copy(sqpb200_ctx *c, void *dst, const void *src, size_t bytes, void *stream) {
    if (!c || (bytes && (!dst || !src))) return SQPB200_ERR_INVALID;
    CK(c, cudaSetDevice(c->device));
    if (bytes) CK(c, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, (cudaStream_t)stream));
    return SQPB200_OK;
}
int sqpb200_stream_create(sqpb200_ctx *c, void **stream) {
    if (!c || !stream) return SQPB200_ERR_INVALID;
    CK(c, cudaSetDevice(c->device));
    cudaStream_t s = nullptr;
    CK(c, cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    *stream = s;
    return SQPB200_OK;
}
int sqpb200_stream_destroy(sqpb200_ctx *c, void *stream) {
    if (!c) return SQPB200_ERR_INVALID;
    CK(c, cudaSetDevice(c->device));
    if (stream) CK(c, cudaStreamDestroy((cudaStream_t)stream));
    return SQPB200_OK;
}
int sqpb200_stream_sync(sqpb200_ctx *c, void *stream) {
    if (!c) return SQPB200_ERR_INVALID;
    CK(c, cudaSetDevice(c->device));
    CK(c, cudaStreamSynchronize((cudaStream_t)stream));
    return SQPB200_OK;
}
int sqpb200_host_alloc(sqpb200_ctx *c, size_t bytes, void **host_ptr) {
    if (!c || !host_ptr) return SQPB200_ERR_INVALID;
    *host_ptr = nullptr;
    CK(c, cudaSetDevice(c->device));
    cudaError_t e = cudaMallocHost(host_ptr, bytes ? bytes : 1);
    if (e != cudaSuccess) return fail(c, SQPB200_ERR_NOMEM, "sqpb200_host_alloc", e);
    return SQPB200_OK;
}
int sqpb200_host_free(sqpb200_ctx *c, void *host_ptr) {
    if (!c) return SQPB200_ERR_INVALID;
    CK(c, cudaSetDevice(c->device));
    if (host_ptr) CK(c, cudaFreeHost(host_ptr));
    return SQPB200_OK;
}
int sqpb200_ipc_export(sqpb200_ctx *c, const void *dev_ptr, unsigned char handle[SQPB200_IPC_HANDLE_BYTES]) {
    if (!c || !dev_ptr || !handle) return SQPB200_ERR_INVALID;
    static_assert(sizeof(cudaIpcMemHandle_t) == SQPB200_IPC_HANDLE_BYTES, "CUDA IPC handle size");
    CK(c, cudaSetDevice(c->device));
    cudaIpcMemHandle_t h;
    CK(c, cudaIpcGetMemHandle(&h, const_cast<void *>(dev_ptr)));
    memcpy(handle, &h, sizeof h);
    return SQPB200_OK;
}
int sqpb200_ipc_import(sqpb200_ctx *c, const unsigned char handle[SQPB200_IPC_HANDLE_BYTES], void **dev_ptr) {
    if (!c || !handle || !dev_ptr) return SQPB200_ERR_INVALID;
    *dev_ptr = nullptr;
    CK(c, cudaSetDevice(c->device));
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof h);
    CK(c, cudaIpcOpenMemHandle(dev_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return SQPB200_OK;
}
int sqpb200_ipc_release(sqpb200_ctx *c, void *dev_ptr) {
    if (!c) return SQPB200_ERR_INVALID;
    CK(c, cudaSetDevice(c->device));
    if (dev_ptr) CK(c, cudaIpcCloseMemHandle(dev_ptr));
    return SQPB200_OK;
}

int sqpb200_qp_batch_set_precision(sqpb200_qp_batch *b, int fp32) {
    if (!b) return SQPB200_ERR_INVALID;
    if (fp32 != 0 && fp32 != 1) return fail(b->ctx, SQPB200_ERR_INVALID, "sqpb200_qp_batch_set_precision: 0 (fp64) or 1 (fp32)");
    if (b->f32 != fp32) {  // a factor stored by the other instantiation is not this one's: forget it (separate solve() and REUSE_FACTOR)
        b->fact_valid = false;
        CK(b->ctx, cudaSetDevice(b->ctx->device));
        CK(b->ctx, cudaDeviceSynchronize());
        CK(b->ctx, cudaMemset(b->fact_rho, 0xff, (size_t)b->batch * sizeof(double)));  // all-ones bit pattern is a NaN: no factor stored
    }
    b->f32 = fp32;
    return SQPB200_OK;
}

int sqpb200_qp_batch_destroy(sqpb200_qp_batch *b) {
    if (!b) return SQPB200_OK;
    cudaSetDevice(b->ctx->device);
    cudaDeviceSynchronize();
    void *ptrs[] = {b->x, b->y, b->z, b->status, b->iter, b->rho_updates, b->rho_estimate, b->res_prim, b->res_dual,
                    b->rho, b->ctype, b->fact, b->fact_rho, b->total_iters, b->dP, b->dq, b->dA, b->dl, b->du, b->sp_outer, b->sp_inner, b->sp_vals, b->sp2_outer, b->sp2_inner, b->sp2_perm, b->sp_pack, b->cl_scratch, b->gen_scratch, b->rq};
    for (void *p : ptrs)
        if (p) cudaFree(p);
    if (b->last_event) cudaEventDestroy(b->last_event);
    if (b->stage_event) cudaEventDestroy(b->stage_event);
    if (b->ready_dev) cudaFree(b->ready_dev);
    if (b->ready_host) cudaFreeHost(b->ready_host);
    delete b;
    return SQPB200_OK;
}

int sqpb200_qp_batch_create(sqpb200_ctx *c, int batch, int n, int m, sqpb200_qp_batch **out) {
    if (!c || !out) return SQPB200_ERR_INVALID;
    *out = nullptr;
    if (batch < 1 || n < 1 || m < 0) return fail(c, SQPB200_ERR_INVALID, "sqpb200_qp_batch_create: need batch >= 1, n >= 1, m >= 0");
    if (!generic_supported(n, m, c->prop.sharedMemPerBlockOptin) && !tile_supported(n, m))
        return fail(c, SQPB200_ERR_UNSUPPORTED, "sqpb200_qp_batch_create: (n, m) too large for the kernels' shared-memory vectors");
    CK(c, cudaSetDevice(c->device));
    sqpb200_qp_batch *b = new (std::nothrow) sqpb200_qp_batch();
    if (!b) return fail(c, SQPB200_ERR_NOMEM, "host allocation failed");
    b->ctx = c;
    if (cudaEventCreateWithFlags(&b->last_event, cudaEventDisableTiming) != cudaSuccess) {
        delete b;
        return fail(c, SQPB200_ERR_CUDA, "sqpb200_qp_batch_create: cudaEventCreate");
    }
    b->batch = batch;
    b->n = n;
    b->m = m;
    size_t B = (size_t)batch;
    size_t mm = m > 0 ? (size_t)m : 1;
    cudaError_t e = cudaSuccess;
    auto alloc = [&](void **p, size_t bytes) {
        if (e == cudaSuccess) e = cudaMalloc(p, bytes);
    };
    alloc((void **)&b->x, B * n * sizeof(double));
    alloc((void **)&b->y, B * mm * sizeof(double));
    alloc((void **)&b->z, B * mm * sizeof(double));
    alloc((void **)&b->status, B * sizeof(int));
    alloc((void **)&b->iter, B * sizeof(int));
    alloc((void **)&b->rho_updates, B * sizeof(int));
    alloc((void **)&b->rho_estimate, B * sizeof(double));
    alloc((void **)&b->res_prim, B * sizeof(double));
    alloc((void **)&b->res_dual, B * sizeof(double));
    alloc((void **)&b->rho, B * sizeof(double));
    alloc((void **)&b->ctype, B * mm);
    alloc((void **)&b->fact_rho, B * sizeof(double));
    alloc((void **)&b->total_iters, sizeof(unsigned long long));
    if (e != cudaSuccess) {
        int rc = fail(c, e == cudaErrorMemoryAllocation ? SQPB200_ERR_NOMEM : SQPB200_ERR_CUDA, "sqpb200_qp_batch_create: cudaMalloc", e);
        sqpb200_qp_batch_destroy(b);
        return rc;
    }
    // B default-constructed solvers: QPSolverInfo defaults, qp.hpp:72-79
    cudaStream_t s = c->stream;
    cudaMemsetAsync(b->x, 0, B * n * sizeof(double), s);
    cudaMemsetAsync(b->y, 0, B * mm * sizeof(double), s);
    cudaMemsetAsync(b->z, 0, B * mm * sizeof(double), s);
    cudaMemsetAsync(b->iter, 0, B * sizeof(int), s);
    cudaMemsetAsync(b->rho_updates, 0, B * sizeof(int), s);
    cudaMemsetAsync(b->rho_estimate, 0, B * sizeof(double), s);
    cudaMemsetAsync(b->res_prim, 0, B * sizeof(double), s);
    cudaMemsetAsync(b->res_dual, 0, B * sizeof(double), s);
    cudaMemsetAsync(b->rho, 0, B * sizeof(double), s);
    cudaMemsetAsync(b->ctype, 0, B * mm, s);
    cudaMemsetAsync(b->fact_rho, 0xff, B * sizeof(double), s);  // all-ones bit pattern is a NaN: no factor stored
    cudaMemsetAsync(b->total_iters, 0, sizeof(unsigned long long), s);
    // status = UNINITIALIZED (4): byte pattern 0x04040404 is not 4, so fill through a tiny kernel-free path
    {
        int *h = (int *)malloc(B * sizeof(int));
        if (!h) {
            sqpb200_qp_batch_destroy(b);
            return fail(c, SQPB200_ERR_NOMEM, "host allocation failed");
        }
        for (size_t i = 0; i < B; ++i) h[i] = SQPB200_UNINITIALIZED;
        e = cudaMemcpyAsync(b->status, h, B * sizeof(int), cudaMemcpyHostToDevice, s);
        cudaStreamSynchronize(s);
        free(h);
    }
    if (e != cudaSuccess || (e = cudaGetLastError()) != cudaSuccess) {
        int rc = fail(c, SQPB200_ERR_CUDA, "sqpb200_qp_batch_create: init", e);
        sqpb200_qp_batch_destroy(b);
        return rc;
    }
    *out = b;
    return SQPB200_OK;
}

}  // extern "C"

// ---- launch plumbing -----------------------------------------------------------------------

enum { KERNEL_NONE = 0, KERNEL_GENERIC, KERNEL_TILE, KERNEL_BLOCK, KERNEL_CLUSTER, KERNEL_SMALL };

static int ensure_scratch(sqpb200_qp_batch *b, size_t bytes) {
    sqpb200_ctx *c = b->ctx;
    if (bytes <= b->gen_scratch_bytes) return SQPB200_OK;
    CK(c, cudaDeviceSynchronize());
    if (b->gen_scratch) cudaFree(b->gen_scratch);
    b->gen_scratch = nullptr;
    b->gen_scratch_bytes = 0;
    cudaError_t e = cudaMalloc(&b->gen_scratch, bytes);
    if (e != cudaSuccess) return fail(c, SQPB200_ERR_NOMEM, "scratch cudaMalloc", e);
    b->gen_scratch_bytes = bytes;
    return SQPB200_OK;
}

// Order `stream` behind the last launch on this object when that ran on another stream (a non-blocking stream does not
// synchronise with the legacy default stream, so results read on stream 0 would otherwise be stale).
static int order_after_last_launch(sqpb200_qp_batch *b, cudaStream_t stream) {
    if (b->last_event && b->last_stream != stream) CK(b->ctx, cudaStreamWaitEvent(stream, b->last_event, 0));
    return SQPB200_OK;
}

static int ensure_fact(sqpb200_qp_batch *b, size_t doubles_per_qp) {
    if (b->fact && b->fact_doubles >= doubles_per_qp) return SQPB200_OK;
    if (b->fact) {
        cudaDeviceSynchronize();
        cudaFree(b->fact);
        b->fact = nullptr;
        b->fact_valid = false;
    }
    b->fact_doubles = doubles_per_qp;
    cudaError_t e = cudaMalloc(&b->fact, sizeof(double) * (size_t)b->batch * doubles_per_qp);
    if (e != cudaSuccess) return fail(b->ctx, SQPB200_ERR_NOMEM, "factor slab cudaMalloc", e);
    return SQPB200_OK;
}

static int ensure_staging(sqpb200_qp_batch *b) {
    if (!b->ready_dev) {
        cudaError_t e0 = cudaMalloc(&b->ready_dev, sizeof(int));
        if (e0 == cudaSuccess) e0 = cudaMallocHost(&b->ready_host, sizeof(int) * 96);
        if (e0 == cudaSuccess) e0 = cudaEventCreateWithFlags(&b->stage_event, cudaEventDisableTiming);
        if (e0 != cudaSuccess) return fail(b->ctx, SQPB200_ERR_NOMEM, "staging flag allocation", e0);
    }
    if (b->dP) return SQPB200_OK;
    size_t B = (size_t)b->batch, n = b->n, m = b->m > 0 ? b->m : 1;
    cudaError_t e = cudaMalloc(&b->dP, B * n * n * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc(&b->dq, B * n * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc(&b->dA, B * m * n * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc(&b->dl, B * m * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc(&b->du, B * m * sizeof(double));
    if (e != cudaSuccess) return fail(b->ctx, SQPB200_ERR_NOMEM, "input staging cudaMalloc", e);
    return SQPB200_OK;
}

// One kernel launch over QPs [first, first+count) of the batch arrays.
// results written by the kernel straight into caller arrays instead of the object's state (sqpb200_qp_batch_setup_solve_to)
struct OutOverride {
    double *x, *y, *z;
    int *status, *iter, *rho_updates;
    double *rho_estimate, *res_prim, *res_dual;
};

static int launch_range(sqpb200_qp_batch *b, const sqpb200_qp_settings *st, unsigned mode, int first, int count,
                        const double *P, const double *q, const double *A, const double *l, const double *u,
                        cudaStream_t stream, const int *ready = nullptr, const SparseA *sp = nullptr, const OutOverride *ov = nullptr,
                        bool *ov_used = nullptr) {
    sqpb200_ctx *c = b->ctx;
    KernelParams p{};
    p.first = first;
    p.count = count;
    p.n = b->n;
    p.m = b->m;
    p.P = P; p.q = q; p.A = A; p.l = l; p.u = u;
    p.x = b->x; p.z = b->z; p.y = b->y;
    p.status = b->status; p.iter = b->iter; p.rho_updates = b->rho_updates;
    p.rho_estimate = b->rho_estimate; p.res_prim = b->res_prim; p.res_dual = b->res_dual; p.rho = b->rho;
    p.ctype = b->ctype;
    p.fact_rho = b->fact_rho;
    p.total_iters = b->total_iters;
    p.mode = mode;
    p.ready = ready;
    p.s = *st;
    if (sp) p.sp = *sp;
    int slot = (int)(c->launches % sqpb200_ctx::kCounters);
    p.work_counter = c->counters + slot;
    CK(c, cudaMemsetAsync(p.work_counter, 0, sizeof(int), stream));

    // kernel choice (dense A): thread-per-QP literal KKT kernel (n + m <= 16: the SQP regime) > register-tiled (n <= 64, m <= 128) >
    // blocked (n <= 256, m <= 1024) > generic (anything that fits). Sparse A: cluster kernel when the instance fits on chip, else blocked.
    const size_t optin = c->prop.sharedMemPerBlockOptin;
    const int opt = c->opt_kernel;
    int kind = KERNEL_NONE, clusters = 0;
    if (sp) {
        // a cluster of 4 CTAs per QP with H^-1 distributed over their shared memory when the instance fits, else the blocked kernel
        if ((opt == 0 || opt == 4) && sp->cluster_size > 0)
            clusters = cluster_max_clusters(b->n, b->m, sp->nnz, sp->col_slice_cap, sp->cluster_size);
        // Eight CTAs per QP buy latency, not throughput (measured at n = 256, nnz = 8.9 k: 0.84 ms per QP on 8 SMs against 7.7 ms on one SM
        // with the blocked kernel, i.e. the same QPs per SM-second): keep them for batches that cannot fill the SMs one QP each
        if (clusters >= 1 && sp->cluster_size == 8 && opt == 0 && count > c->prop.multiProcessorCount &&
            block_sparse_supported(b->n, b->m, sp->nnz, optin))
            clusters = 0;
        if (opt == 4 && clusters < 1) return fail(c, SQPB200_ERR_UNSUPPORTED, "cluster kernel forced but the problem is outside its range");
        kind = clusters >= 1 ? KERNEL_CLUSTER : KERNEL_BLOCK;
        // solve() after setup()/update_qp(): the blocked kernel's factor lives in the slab, the cluster kernel rebuilds its own --
        // keep the kernel that ran the setup
        if ((mode & MODE_LOAD_FACTOR) && b->fact_valid) {
            if (b->fact_kernel == KERNEL_BLOCK) kind = KERNEL_BLOCK, clusters = 0;
            else if (b->fact_kernel == KERNEL_CLUSTER && clusters < 1)
                clusters = cluster_max_clusters(b->n, b->m, sp->nnz, sp->col_slice_cap, sp->cluster_size), kind = KERNEL_CLUSTER;
        }
        if (kind == KERNEL_CLUSTER && clusters < 1) return fail(c, SQPB200_ERR_UNSUPPORTED, "cluster kernel required by the stored state but unavailable");
    } else if (opt == 1) {
        kind = KERNEL_GENERIC;
    } else if (opt == 2) {
        if (!tile_supported(b->n, b->m)) return fail(c, SQPB200_ERR_UNSUPPORTED, "register-tiled kernel forced but (n, m) is outside its range");
        kind = KERNEL_TILE;
    } else if (opt == 3) {
        if (!block_supported(b->n, b->m, optin)) return fail(c, SQPB200_ERR_UNSUPPORTED, "blocked kernel forced but (n, m) is outside its range");
        kind = KERNEL_BLOCK;
    } else if (opt == 5) {
        if (!small_supported(b->n, b->m)) return fail(c, SQPB200_ERR_UNSUPPORTED, "thread-per-QP kernel forced but n + m > 16");
        kind = KERNEL_SMALL;
    } else if (opt == 4) {
        return fail(c, SQPB200_ERR_UNSUPPORTED, "cluster kernel forced but A is dense");
    } else if (st->verbose && generic_supported(b->n, b->m, optin)) {
        // settings.verbose (the reference's per-check status line, qp.cpp:114-118 / :375-382) is a debugging aid: such calls take the
        // generic kernel, the one that prints it, so the fast kernels carry no printf in their loops
        kind = KERNEL_GENERIC;
    } else if (small_supported(b->n, b->m)) {
        kind = KERNEL_SMALL;
    } else if (tile_supported(b->n, b->m)) {
        kind = KERNEL_TILE;
    } else if (b->f32 && generic_supported(b->n, b->m, optin)) {
        kind = KERNEL_GENERIC;  // QPSolver<float> beyond the tile kernel's shapes: the generic kernel's float instantiation
    } else if (block_supported(b->n, b->m, optin)) {
        kind = KERNEL_BLOCK;
    } else {
        kind = KERNEL_GENERIC;
    }
    // The layout of a stored factor belongs to the kernel that wrote it (dense H^-1 for the tile and generic kernels, blocked
    // LDL^T panels for the blocked kernel, nothing for the thread-per-QP kernel): solve() after setup()/update_qp() keeps that
    // kernel even when the option or settings.verbose changed in between, and a REUSE of a factor another kernel kept is dropped.
    if ((mode & MODE_LOAD_FACTOR) && b->fact_valid && b->fact_kernel != KERNEL_NONE && !sp) {
        // the fp64 tile and generic kernels both store the dense n x n inverse (fp32: the tile kernel widens to double, the generic
        // kernel packs floats)
        const bool compatible = (kind == b->fact_kernel) || (!b->f32 && ((kind == KERNEL_TILE && b->fact_kernel == KERNEL_GENERIC) ||
                                                                         (kind == KERNEL_GENERIC && b->fact_kernel == KERNEL_TILE)));
        if (!compatible) kind = b->fact_kernel;
    }
    if ((mode & MODE_REUSE) && b->keep_kernel != kind) mode &= ~MODE_REUSE;
    if (ov_used) *ov_used = false;
    if (ov && kind == KERNEL_TILE) {  // the register-tiled kernel honours MODE_FRESH: its epilogue writes the caller's arrays
        mode |= MODE_FRESH;
        p.x = ov->x; p.y = ov->y; p.z = ov->z;
        p.status = ov->status; p.iter = ov->iter; p.rho_updates = ov->rho_updates;
        p.rho_estimate = ov->rho_estimate; p.res_prim = ov->res_prim; p.res_dual = ov->res_dual;
        if (ov_used) *ov_used = true;
    } else if (ov) {
        // the other kernels read-modify-write the object's info arrays: make them those of default-constructed solvers
        // (rho_updates 0, ...; status is overwritten by the factorisation) so that the caller sees per-call values, as documented
        const size_t B = (size_t)count;
        CK(c, cudaMemsetAsync(b->rho_updates, 0, B * sizeof(int), stream));
        CK(c, cudaMemsetAsync(b->iter, 0, B * sizeof(int), stream));
        CK(c, cudaMemsetAsync(b->rho_estimate, 0, B * sizeof(double), stream));
        CK(c, cudaMemsetAsync(b->res_prim, 0, B * sizeof(double), stream));
        CK(c, cudaMemsetAsync(b->res_dual, 0, B * sizeof(double), stream));
    }
    p.mode = mode;
    if (mode & MODE_KEEP_INITIAL) b->keep_kernel = kind;
    if (mode & MODE_STORE_FACTOR) b->fact_kernel = kind;
    // Time slicing (register-tiled kernel, fused launches of fresh instances): with only a few QPs per resident CTA -- one batch split
    // over several GPUs -- whole 500...1001-iteration solves leave a long tail when the queue runs dry (list scheduling: about half a QP
    // per CTA slot); slices of 250 iterations that re-enter the queue cut it to a fraction. Results are bit-identical to an unsliced solve.
    int slice = 0;
    if (kind == KERNEL_TILE && (mode & ~MODE_FRESH) == (MODE_RESET | MODE_FACTOR | MODE_SOLVE) && !ready &&
        tile_sliceable(b->n, b->m, c->opt_tile_warps, b->f32) && st->max_iter > 0) {
        const int slots = tile_slots(c->prop.multiProcessorCount);
        int slice_first = 0;
        if (c->opt_slice > 0) slice = c->opt_slice & 0xffff, slice_first = c->opt_slice >> 16;
        // measured (1 B200, S1, inputs in local HBM): 1024 QPs 3.25 ms unsliced, 3.14 ms in slices of 250, 3.09 ms with a first slice of
        // 500 iterations (every solve of this workload needs them anyway) and slices of 125 after it; 2048 QPs 6.03 -> 6.10 ms: automatic
        // below ~4 QPs per CTA slot only. With the inputs in PEER memory and the results written to caller arrays (the multi-GPU flow,
        // 1024 QPs per GPU; the first slice leaves local copies of P, A, q, l, u behind so that a resume reads nothing over NVLink) it
        // gains with one remote GPU (two GPUs: 3.41 -> 3.24 ms) but loses with seven of them pulling from the owner (eight GPUs: 3.47 ->
        // 3.54 ... 3.62 ms: all first slices, and with them all transfers, crowd into the first half of the launch), so the automatic mode
        // leaves those launches alone. (value > 0: iterations per slice + 65536 x iterations of the first slice.)
        else if (c->opt_slice < 0 && !ov && count <= 4 * slots && count > slots / 4 && st->max_iter >= 500)
            slice = st->max_iter / 8, slice_first = st->max_iter / 2;
        if (slice >= st->max_iter) slice = 0;
        p.slice_first = slice_first > slice ? slice_first : slice;
    }
    if (slice > 0) {
        const int cap = count * ((st->max_iter + slice - 1) / slice + 1);
        if (cap > b->rq_cap) {
            CK(c, cudaDeviceSynchronize());
            if (b->rq) cudaFree(b->rq);
            b->rq = nullptr;
            b->rq_cap = 0;
            cudaError_t ea = cudaMalloc(&b->rq, sizeof(int) * ((size_t)cap + 2));
            if (ea != cudaSuccess) return fail(c, SQPB200_ERR_NOMEM, "re-queue ring cudaMalloc", ea);
            b->rq_cap = cap;
        }
        CK(c, cudaMemsetAsync(b->rq, 0xff, sizeof(int) * (size_t)b->rq_cap, stream));
        CK(c, cudaMemsetAsync(b->rq + b->rq_cap, 0, 2 * sizeof(int), stream));
        p.slice_iters = slice;
        p.rq = b->rq;
        p.rq_cap = b->rq_cap;
        p.rq_alloc = b->rq + b->rq_cap;
        p.done = b->rq + b->rq_cap + 1;
        p.sus_x = b->x; p.sus_z = b->z; p.sus_y = b->y;
        p.sus_status = b->status; p.sus_iter = b->iter; p.sus_rho_updates = b->rho_updates;
        p.sus_rho_estimate = b->rho_estimate; p.sus_res_prim = b->res_prim; p.sus_res_dual = b->res_dual;
        if (P != b->dP) {  // inputs that are not the object's own staging copies may live in peer memory: keep local copies for the resumes
            int rc = ensure_staging(b);
            if (rc) return rc;
            p.loc_P = b->dP;
            p.loc_A = b->dA;
            p.loc_q = b->dq;
            p.loc_l = b->dl;
            p.loc_u = b->du;
        }
    }
    const bool needs_fact = kind == KERNEL_BLOCK || kind == KERNEL_GENERIC || slice > 0 ||
                            (kind == KERNEL_TILE && (mode & (MODE_STORE_FACTOR | MODE_LOAD_FACTOR | MODE_KEEP_INITIAL | MODE_REUSE)));
    if (needs_fact) {
        int rc = ensure_fact(b, kind == KERNEL_BLOCK ? block_fact_doubles(b->n) : (size_t)b->n * b->n);
        if (rc) return rc;
    }
    p.fact = b->fact;
    cudaError_t e;
    if (kind == KERNEL_CLUSTER) {
        if (clusters > count) clusters = count;
        const size_t need = cluster_scratch_bytes(cluster_max_clusters(b->n, b->m, sp->nnz, sp->col_slice_cap, sp->cluster_size));
        if (need > b->cl_scratch_bytes) {
            CK(c, cudaDeviceSynchronize());
            if (b->cl_scratch) cudaFree(b->cl_scratch);
            b->cl_scratch = nullptr;
            b->cl_scratch_bytes = 0;
            cudaError_t ea = cudaMalloc(&b->cl_scratch, need);
            if (ea != cudaSuccess) return fail(c, SQPB200_ERR_NOMEM, "cluster scratch cudaMalloc", ea);
            b->cl_scratch_bytes = need;
        }
        e = launch_cluster(p, clusters, b->cl_scratch, stream, c->last_kernel, sizeof c->last_kernel);
    } else if (kind == KERNEL_SMALL) {
        e = launch_small(p, b->f32, stream, c->last_kernel, sizeof c->last_kernel);
    } else if (kind == KERNEL_TILE) {
        e = launch_tile(p, c->prop.multiProcessorCount, c->opt_ctas_per_sm, c->opt_tile_warps, b->f32, stream, c->last_kernel, sizeof c->last_kernel);
    } else if (kind == KERNEL_BLOCK) {
        e = launch_block(p, c->prop.multiProcessorCount, optin, stream, c->last_kernel, sizeof c->last_kernel);
    } else {
        if (!generic_supported(b->n, b->m, optin)) return fail(c, SQPB200_ERR_UNSUPPORTED, "(n, m) too large for the generic kernel");
        int grid = generic_grid(count, c->prop.multiProcessorCount);
        int rc = ensure_scratch(b, generic_scratch_bytes(b->n, grid));
        if (rc) return rc;
        p.scratch = b->gen_scratch;
        e = launch_generic(p, c->prop.multiProcessorCount, optin, b->f32, stream, nullptr);
        snprintf(c->last_kernel, sizeof c->last_kernel, b->f32 ? "generic<f32>" : "generic");
    }
    if (e != cudaSuccess) return fail(c, SQPB200_ERR_CUDA, "kernel launch", e);
    c->launches += 1;
    b->last_stream = stream;
    CK(c, cudaEventRecord(b->last_event, stream));
    return SQPB200_OK;
}

// Chunk boundaries of a staged HOST_PTRS call. The transfer delivers QPs only ~1.5x faster than the resident CTAs consume them, so for
// the first few waves every CTA that finishes a QP is waiting for the next flag: a geometric ramp (an eighth, a quarter, a half of a wave
// of resident CTAs, a wave), then half-wave steps for eight waves, and only then the equal chunks (count / chunks) that keep the number
// of copy calls down. bound[0] = 0 < bound[1] < ... < bound[return value] = count; at most cap - 1 chunks.
static int chunk_bounds(int count, int chunks, int wave, int *bound, int cap) {
    int nb = 0;
    bound[nb++] = 0;
    const int uniform = count / chunks > 0 ? count / chunks : 1;
    const int room = cap - chunks - 2;  // entries the fine-grained prefix may use
    int r = wave / 8 > 0 ? wave / 8 : 1;
    for (; r < wave && r < uniform && nb < room; r *= 2) bound[nb++] = r;
    const int half = wave / 2 > 0 ? wave / 2 : 1;
    for (r = wave; half < uniform && r < 8 * wave && r < count && nb < room; r += half) bound[nb++] = r;
    for (int k = 1; k <= chunks; ++k) {
        const int hi = (int)((size_t)count * k / chunks);
        if (hi > bound[nb - 1]) bound[nb++] = hi;
    }
    return nb - 1;
}

static int run(sqpb200_qp_batch *b, const sqpb200_qp_settings *st, unsigned mode, int count, const double *P,
               const double *q, const double *A, const double *l, const double *u, unsigned flags, void *stream_) {
    if (!b) return SQPB200_ERR_INVALID;
    sqpb200_ctx *c = b->ctx;
    if (!st) return fail(c, SQPB200_ERR_INVALID, "settings is NULL");
    if (count < 0 || count > b->batch) return fail(c, SQPB200_ERR_INVALID, "count outside [0, batch]");
    if (count == 0) return SQPB200_OK;
    if (!P || !q || !A || ((!l || !u) && b->m > 0)) return fail(c, SQPB200_ERR_INVALID, "NULL problem array");
    CK(c, cudaSetDevice(c->device));
    cudaStream_t stream = (cudaStream_t)stream_;  // NULL is the CUDA legacy default stream
    CK(c, cudaMemsetAsync(b->total_iters, 0, sizeof(unsigned long long), stream));
    const size_t n = b->n, m = b->m;
    if (flags & SQPB200_DEVICE_PTRS) return launch_range(b, st, mode, 0, count, P, q, A, l, u, stream);

    // HOST_PTRS: ONE persistent launch over the whole batch; the inputs are staged chunk by chunk on the copy
    // stream while it runs. After each chunk a 4-byte copy publishes how many QPs have landed; a CTA that
    // draws a QP beyond that count waits (draw_qp). H2D and compute overlap with no per-chunk launch tails.
    int rc = ensure_staging(b);
    if (rc) return rc;
    // chunking hides the transfer behind the solve; below ~4 MB per chunk the copy calls cost more than they hide (an SQP outer
    // iteration hands over a few hundred KB): fewer, larger chunks then
    const size_t total_bytes = (size_t)count * (n * n + n + m * n + 2 * m) * sizeof(double);
    int chunks = c->opt_chunks;
    if ((size_t)chunks > total_bytes / (4u << 20) + 1) chunks = (int)(total_bytes / (4u << 20) + 1);
    if (chunks > count) chunks = count;
    if (chunks <= 1 || (flags & SQPB200_HOST_ASYNC)) {
        // small call: the copies go on the caller's stream ahead of the launch -- no copy stream, no flag, no cross-stream event
        CK(c, cudaMemcpyAsync(b->dP, P, (size_t)count * n * n * sizeof(double), cudaMemcpyHostToDevice, stream));
        CK(c, cudaMemcpyAsync(b->dA, A, (size_t)count * m * n * sizeof(double), cudaMemcpyHostToDevice, stream));
        CK(c, cudaMemcpyAsync(b->dq, q, (size_t)count * n * sizeof(double), cudaMemcpyHostToDevice, stream));
        if (m > 0) {
            CK(c, cudaMemcpyAsync(b->dl, l, (size_t)count * m * sizeof(double), cudaMemcpyHostToDevice, stream));
            CK(c, cudaMemcpyAsync(b->du, u, (size_t)count * m * sizeof(double), cudaMemcpyHostToDevice, stream));
        }
        rc = launch_range(b, st, mode, 0, count, b->dP, b->dq, b->dA, b->dl, b->du, stream);
        if (rc) return rc;
        if (!(flags & SQPB200_HOST_ASYNC)) CK(c, cudaStreamSynchronize(stream));
        return SQPB200_OK;
    }
    CK(c, cudaMemsetAsync(b->ready_dev, 0, sizeof(int), stream));
    // the compute stream may still be reading the staging buffers from an earlier call; the flag reset must precede the copies
    CK(c, cudaEventRecord(b->stage_event, stream));
    CK(c, cudaStreamWaitEvent(c->copy_stream, b->stage_event, 0));
    int bound[96];
    chunks = chunk_bounds(count, chunks, 2 * c->prop.multiProcessorCount, bound, 96);
    for (int k = 0; k < chunks; ++k) {
        size_t lo = (size_t)bound[k], hi = (size_t)bound[k + 1], cnt = hi - lo;
        CK(c, cudaMemcpyAsync(b->dP + lo * n * n, P + lo * n * n, cnt * n * n * sizeof(double), cudaMemcpyHostToDevice, c->copy_stream));
        CK(c, cudaMemcpyAsync(b->dA + lo * m * n, A + lo * m * n, cnt * m * n * sizeof(double), cudaMemcpyHostToDevice, c->copy_stream));
        CK(c, cudaMemcpyAsync(b->dq + lo * n, q + lo * n, cnt * n * sizeof(double), cudaMemcpyHostToDevice, c->copy_stream));
        if (m > 0) {
            CK(c, cudaMemcpyAsync(b->dl + lo * m, l + lo * m, cnt * m * sizeof(double), cudaMemcpyHostToDevice, c->copy_stream));
            CK(c, cudaMemcpyAsync(b->du + lo * m, u + lo * m, cnt * m * sizeof(double), cudaMemcpyHostToDevice, c->copy_stream));
        }
        b->ready_host[k] = (int)hi;
        CK(c, cudaMemcpyAsync(b->ready_dev, b->ready_host + k, sizeof(int), cudaMemcpyHostToDevice, c->copy_stream));
    }
    // (the launch follows the copy enqueues on purpose: launching first measured 0.05 ms faster, but under a tool that serialises kernel
    // launches -- ncu, compute-sanitizer, CUDA_LAUNCH_BLOCKING -- the kernel would then wait for copies the host has not issued yet)
    rc = launch_range(b, st, mode, 0, count, b->dP, b->dq, b->dA, b->dl, b->du, stream, b->ready_dev);
    if (rc) {
        cudaStreamSynchronize(c->copy_stream);
        return rc;
    }
    CK(c, cudaStreamSynchronize(stream));
    return SQPB200_OK;
}

extern "C" {

int sqpb200_staging_chunk_bounds(int count, int chunks, int wave, int *bound, int cap) {
    if (count < 1 || chunks < 1 || wave < 1 || !bound || cap < chunks + 3) return -1;
    if (chunks > count) chunks = count;
    return chunk_bounds(count, chunks, wave, bound, cap);
}

int sqpb200_qp_batch_setup(sqpb200_qp_batch *b, const sqpb200_qp_settings *s, int count, const double *P, const double *q,
                           const double *A, const double *l, const double *u, unsigned flags, void *stream) {
    int rc = run(b, s, MODE_RESET | MODE_FACTOR | MODE_STORE_FACTOR, count, P, q, A, l, u, flags, stream);
    if (!rc && b) b->fact_valid = true;
    return rc;
}
int sqpb200_qp_batch_update_qp(sqpb200_qp_batch *b, const sqpb200_qp_settings *s, int count, const double *P,
                               const double *q, const double *A, const double *l, const double *u, unsigned flags, void *stream) {
    int rc = run(b, s, MODE_FACTOR | MODE_STORE_FACTOR, count, P, q, A, l, u, flags, stream);
    if (!rc && b) b->fact_valid = true;
    return rc;
}
int sqpb200_qp_batch_solve(sqpb200_qp_batch *b, const sqpb200_qp_settings *s, int count, const double *P, const double *q,
                           const double *A, const double *l, const double *u, unsigned flags, void *stream) {
    // Without a stored factor every instance is still UNINITIALIZED (solve is a no-op, qp.cpp:68-71)
    // or was last run through the fused setup_solve, which keeps the factor on-chip only.
    if (b && !b->fact_valid) {
        bool fresh = true;  // never set up at all: run the launch so the no-op semantics are exercised on device
        if (b->fused_used) fresh = false;
        if (!fresh)
            return fail(b->ctx, SQPB200_ERR_INVALID,
                        "sqpb200_qp_batch_solve: the last setup was the fused setup_solve, which does not keep the factor; call setup() first");
    }
    return run(b, s, MODE_LOAD_FACTOR | MODE_SOLVE | MODE_STORE_FACTOR, count, P, q, A, l, u, flags, stream);
}
int sqpb200_qp_batch_setup_solve(sqpb200_qp_batch *b, const sqpb200_qp_settings *s, int count, const double *P,
                                 const double *q, const double *A, const double *l, const double *u, unsigned flags, void *stream) {
    if (b) {
        b->fact_valid = false;
        b->fused_used = true;
    }
    return run(b, s, MODE_RESET | MODE_FACTOR | MODE_SOLVE, count, P, q, A, l, u, flags, stream);
}

int sqpb200_qp_batch_setup_solve_to(sqpb200_qp_batch *b, const sqpb200_qp_settings *s, int count, const double *P, const double *q,
                                    const double *A, const double *l, const double *u, double *x, double *y, double *z, int *status,
                                    int *iter, int *rho_updates, double *rho_estimate, double *res_prim, double *res_dual, void *stream_) {
    if (!b) return SQPB200_ERR_INVALID;
    sqpb200_ctx *c = b->ctx;
    if (!s) return fail(c, SQPB200_ERR_INVALID, "settings is NULL");
    if (count < 0 || count > b->batch) return fail(c, SQPB200_ERR_INVALID, "count outside [0, batch]");
    if (!P || !q || !A || ((!l || !u) && b->m > 0)) return fail(c, SQPB200_ERR_INVALID, "NULL problem array");
    if (!x || !y || !z || !status || !iter || !rho_updates || !rho_estimate || !res_prim || !res_dual)
        return fail(c, SQPB200_ERR_INVALID, "sqpb200_qp_batch_setup_solve_to: every result array must be given");
    if (count == 0) return SQPB200_OK;
    CK(c, cudaSetDevice(c->device));
    cudaStream_t stream = (cudaStream_t)stream_;
    CK(c, cudaMemsetAsync(b->total_iters, 0, sizeof(unsigned long long), stream));
    b->fact_valid = false;
    b->fused_used = true;
    const OutOverride ov{x, y, z, status, iter, rho_updates, rho_estimate, res_prim, res_dual};
    bool direct = false;
    int rc = launch_range(b, s, MODE_RESET | MODE_FACTOR | MODE_SOLVE, 0, count, P, q, A, l, u, stream, nullptr, nullptr, &ov, &direct);
    if (rc || direct) return rc;
    // kernels without the direct epilogue: the object's state holds the results; copy them out behind the launch
    return sqpb200_qp_batch_get(b, count, x, y, z, status, iter, rho_updates, rho_estimate, res_prim, res_dual, SQPB200_DEVICE_PTRS, stream_);
}

int sqpb200_qp_batch_setup_solve_opts(sqpb200_qp_batch *b, const sqpb200_qp_settings *s, int count, const double *P,
                                      const double *q, const double *A, const double *l, const double *u, unsigned flags,
                                      void *stream, unsigned opts) {
    if (!b) return SQPB200_ERR_INVALID;
    if (opts & ~(SQPB200_KEEP_FACTOR | SQPB200_REUSE_FACTOR)) return fail(b->ctx, SQPB200_ERR_INVALID, "unknown opts bit");
    unsigned mode = MODE_RESET | MODE_FACTOR | MODE_SOLVE;
    if (opts & SQPB200_KEEP_FACTOR) mode |= MODE_KEEP_INITIAL;
    if (opts & SQPB200_REUSE_FACTOR) mode |= MODE_REUSE;
    b->fact_valid = false;
    b->fused_used = true;
    return run(b, s, mode, count, P, q, A, l, u, flags, stream);
}

}  // extern "C"

// One call of the sparse-A entry points: `mode` as in the dense ones (setup = RESET|FACTOR|STORE_FACTOR, ...)
static int run_sparse(sqpb200_qp_batch *b, const sqpb200_qp_settings *s, unsigned mode, int count, const double *P,
                      const double *q, const double *A_values, const int *A_outer, const int *A_inner, int nnz,
                      int layout, const double *l, const double *u, unsigned flags, void *stream_) {
    if (!b) return SQPB200_ERR_INVALID;
    sqpb200_ctx *c = b->ctx;
    if (!s) return fail(c, SQPB200_ERR_INVALID, "settings is NULL");
    if (count < 0 || count > b->batch) return fail(c, SQPB200_ERR_INVALID, "count outside [0, batch]");
    if (layout != SQPB200_SPARSE_CSC && layout != SQPB200_SPARSE_CSR) return fail(c, SQPB200_ERR_INVALID, "layout must be SQPB200_SPARSE_CSC or SQPB200_SPARSE_CSR");
    if (nnz < 0 || (nnz > 0 && (!A_values || !A_inner)) || !A_outer || !P || !q || ((!l || !u) && b->m > 0))
        return fail(c, SQPB200_ERR_INVALID, "NULL problem array");
    if (count == 0) return SQPB200_OK;
    CK(c, cudaSetDevice(c->device));
    cudaStream_t stream = (cudaStream_t)stream_;
    const size_t n = b->n, m = b->m, B = count;
    const bool csr = layout == SQPB200_SPARSE_CSR;
    const int n_outer = csr ? b->m : b->n, n_innerdim = csr ? b->n : b->m;
    int rc = ensure_staging(b);
    if (rc) return rc;
    const bool dev = (flags & SQPB200_DEVICE_PTRS) != 0;

    // host copy of the (small, batch-shared) pattern: validated here, and converted to the other compressed view
    std::vector<int> h_outer(n_outer + 1), h_inner(nnz > 0 ? nnz : 1);
    if (dev) {
        CK(c, cudaMemcpyAsync(h_outer.data(), A_outer, sizeof(int) * (n_outer + 1), cudaMemcpyDeviceToHost, stream));
        if (nnz > 0) CK(c, cudaMemcpyAsync(h_inner.data(), A_inner, sizeof(int) * nnz, cudaMemcpyDeviceToHost, stream));
        CK(c, cudaStreamSynchronize(stream));
    } else {
        memcpy(h_outer.data(), A_outer, sizeof(int) * (n_outer + 1));
        if (nnz > 0) memcpy(h_inner.data(), A_inner, sizeof(int) * nnz);
    }
    if (h_outer[0] != 0 || h_outer[n_outer] != nnz) return fail(c, SQPB200_ERR_INVALID, "sparse pattern: outer[0] must be 0 and outer[last] must be nnz");
    for (int o = 0; o < n_outer; ++o)
        if (h_outer[o] > h_outer[o + 1]) return fail(c, SQPB200_ERR_INVALID, "sparse pattern: outer pointers must be non-decreasing");
    for (int e = 0; e < nnz; ++e)
        if (h_inner[e] < 0 || h_inner[e] >= n_innerdim) return fail(c, SQPB200_ERR_INVALID, "sparse pattern: inner index out of range");

    // device copies of the pattern, values and (for host callers) the dense vectors
    if (nnz > b->sp_nnz_cap || !b->sp_outer) {
        CK(c, cudaStreamSynchronize(stream));
        int **ip[] = {&b->sp_outer, &b->sp_inner, &b->sp2_outer, &b->sp2_inner, &b->sp2_perm};
        for (int **q2 : ip) {
            if (*q2) cudaFree(*q2);
            *q2 = nullptr;
        }
        if (b->sp_vals) cudaFree(b->sp_vals);
        b->sp_vals = nullptr;
        if (b->sp_pack) cudaFree(b->sp_pack);
        b->sp_pack = nullptr;
        b->sp_hash = 0;
        const size_t cap = nnz > 0 ? nnz : 1, od = (size_t)(b->m > b->n ? b->m : b->n) + 1;
        cudaError_t e = cudaMalloc(&b->sp_outer, sizeof(int) * od);
        if (e == cudaSuccess) e = cudaMalloc(&b->sp2_outer, sizeof(int) * od);
        if (e == cudaSuccess) e = cudaMalloc(&b->sp_inner, sizeof(int) * cap);
        if (e == cudaSuccess) e = cudaMalloc(&b->sp2_inner, sizeof(int) * cap);
        if (e == cudaSuccess) e = cudaMalloc(&b->sp2_perm, sizeof(int) * cap);
        if (e == cudaSuccess) e = cudaMalloc(&b->sp_vals, sizeof(double) * cap * b->batch);
        if (e == cudaSuccess) e = cudaMalloc(&b->sp_pack, sizeof(unsigned) * 2 * cap);
        if (e != cudaSuccess) return fail(c, SQPB200_ERR_NOMEM, "sparse staging cudaMalloc", e);
        b->sp_nnz_cap = (int)cap;
    }
    const int *d_outer = A_outer, *d_inner = A_inner;
    const double *d_vals = A_values, *dP = P, *dq = q, *dl = l, *du = u;
    if (!dev) {
        CK(c, cudaMemcpyAsync(b->sp_outer, A_outer, sizeof(int) * (n_outer + 1), cudaMemcpyHostToDevice, stream));
        if (nnz > 0) CK(c, cudaMemcpyAsync(b->sp_inner, A_inner, sizeof(int) * nnz, cudaMemcpyHostToDevice, stream));
        // the per-instance arrays (values, P, q, l, u) follow below: chunk by chunk behind the running kernel when a sparse
        // kernel takes the call, in one piece on `stream` when the problem is densified
        d_outer = b->sp_outer; d_inner = b->sp_inner; d_vals = b->sp_vals;
        dP = b->dP; dq = b->dq; dl = b->dl; du = b->du;
    }
    CK(c, cudaMemsetAsync(b->total_iters, 0, sizeof(unsigned long long), stream));

    // Shapes the register-tiled kernel covers keep A in registers anyway: densify. Larger ones run the blocked kernel with the
    // values of one instance staged in shared memory and both compressed views of the pattern.
    // cluster kernel: how many CTAs per QP (0 = does not fit on chip) and the most stored entries in the column slice one CTA owns
    int col_slice_cap = 0, cluster_size = 0;
    if (b->n > 64 && b->n <= 256) {
        std::vector<int> colcount(b->n, 0);
        if (csr) for (int e = 0; e < nnz; ++e) colcount[h_inner[e]]++;
        else for (int j = 0; j < b->n; ++j) colcount[j] = h_outer[j + 1] - h_outer[j];
        cluster_size = cluster_plan(b->n, b->m, nnz, colcount.data(), c->prop.sharedMemPerBlockOptin, &col_slice_cap);
    }
    const bool sparse_kernel = !(s->verbose && c->opt_kernel == 0) && (c->opt_kernel == 0 || c->opt_kernel == 3 || c->opt_kernel == 4) && !(c->opt_kernel == 0 && tile_supported(b->n, b->m)) &&
                               m > 0 && (block_sparse_supported(b->n, b->m, nnz, c->prop.sharedMemPerBlockOptin) || cluster_size > 0);
    auto copy_instances = [&](cudaStream_t cs, size_t lo, size_t cnt) -> cudaError_t {
        cudaError_t e = cudaSuccess;
        if (nnz > 0) e = cudaMemcpyAsync(b->sp_vals + lo * nnz, A_values + lo * nnz, sizeof(double) * cnt * nnz, cudaMemcpyHostToDevice, cs);
        if (e == cudaSuccess) e = cudaMemcpyAsync(b->dP + lo * n * n, P + lo * n * n, sizeof(double) * cnt * n * n, cudaMemcpyHostToDevice, cs);
        if (e == cudaSuccess) e = cudaMemcpyAsync(b->dq + lo * n, q + lo * n, sizeof(double) * cnt * n, cudaMemcpyHostToDevice, cs);
        if (e == cudaSuccess && m > 0) e = cudaMemcpyAsync(b->dl + lo * m, l + lo * m, sizeof(double) * cnt * m, cudaMemcpyHostToDevice, cs);
        if (e == cudaSuccess && m > 0) e = cudaMemcpyAsync(b->du + lo * m, u + lo * m, sizeof(double) * cnt * m, cudaMemcpyHostToDevice, cs);
        return e;
    };
    if (!dev && !sparse_kernel) CK(c, copy_instances(stream, 0, B));
    if (!sparse_kernel) {
        cudaError_t e = launch_densify(d_vals, d_outer, d_inner, nnz, b->m, b->n, csr, count, b->dA, stream);
        if (e != cudaSuccess) return fail(c, SQPB200_ERR_CUDA, "densify launch", e);
        c->launches += nnz > 0 ? 1 : 0;
        rc = launch_range(b, s, mode, 0, count, dP, dq, b->dA, dl, du, stream);
    } else {
        // The derived views of the pattern (the other compressed view, the packed entries) stay on the device between calls with
        // the same pattern (an SQP loop re-solves with new values only): rebuilt only when the pattern's hash changes.
        unsigned long long h = 1469598103934665603ull;
        auto mix = [&h](unsigned v) { h = (h ^ v) * 1099511628211ull; };
        mix((unsigned)layout); mix((unsigned)n_outer); mix((unsigned)nnz);
        for (int o = 0; o <= n_outer; ++o) mix((unsigned)h_outer[o]);
        for (int e = 0; e < nnz; ++e) mix((unsigned)h_inner[e]);
        if (h == 0) h = 1;
        if (h != b->sp_hash) {
        // the other compressed view by a counting sort over the inner index; perm maps its positions to the given order
        std::vector<int> o2(n_innerdim + 1, 0), i2(nnz > 0 ? nnz : 1), perm(nnz > 0 ? nnz : 1);
        for (int e = 0; e < nnz; ++e) o2[h_inner[e] + 1]++;
        for (int k = 0; k < n_innerdim; ++k) o2[k + 1] += o2[k];
        std::vector<int> cursor(o2.begin(), o2.end() - 1);
        for (int o = 0; o < n_outer; ++o)
            for (int e = h_outer[o]; e < h_outer[o + 1]; ++e) {
                const int dst = cursor[h_inner[e]]++;
                i2[dst] = o;
                perm[dst] = e;
            }
        CK(c, cudaMemcpyAsync(b->sp2_outer, o2.data(), sizeof(int) * (n_innerdim + 1), cudaMemcpyHostToDevice, stream));
        if (nnz > 0) {
            CK(c, cudaMemcpyAsync(b->sp2_inner, i2.data(), sizeof(int) * nnz, cudaMemcpyHostToDevice, stream));
            CK(c, cudaMemcpyAsync(b->sp2_perm, perm.data(), sizeof(int) * nnz, cudaMemcpyHostToDevice, stream));
        }
        // packed entries for the cluster kernel: inner index (PACK_BITS bits) | position of the value in the caller's order
        std::vector<unsigned> pack(2 * (size_t)(nnz > 0 ? nnz : 1));
        {
            unsigned *pc = pack.data(), *pr = pack.data() + nnz;  // CSC view, CSR view
            for (int e = 0; e < nnz; ++e) {
                const unsigned given = (unsigned)h_inner[e] | ((unsigned)e << PACK_BITS);
                const unsigned other = (unsigned)i2[e] | ((unsigned)perm[e] << PACK_BITS);
                (csr ? pr : pc)[e] = given;
                (csr ? pc : pr)[e] = other;
            }
        }
        if (nnz > 0) CK(c, cudaMemcpyAsync(b->sp_pack, pack.data(), sizeof(unsigned) * 2 * nnz, cudaMemcpyHostToDevice, stream));
        CK(c, cudaStreamSynchronize(stream));  // the host vectors above go out of scope
        b->sp_hash = h;
        }
        SparseA sp{};
        sp.col_slice_cap = col_slice_cap;
        sp.cluster_size = cluster_size;
        sp.col_pack = b->sp_pack;
        sp.row_pack = b->sp_pack + nnz;
        sp.vals = d_vals;
        sp.nnz = nnz;
        if (csr) {
            sp.row_outer = d_outer; sp.row_inner = d_inner; sp.row_perm = nullptr;
            sp.col_outer = b->sp2_outer; sp.col_inner = b->sp2_inner; sp.col_perm = b->sp2_perm;
        } else {
            sp.col_outer = d_outer; sp.col_inner = d_inner; sp.col_perm = nullptr;
            sp.row_outer = b->sp2_outer; sp.row_inner = b->sp2_inner; sp.row_perm = b->sp2_perm;
        }
        const int *ready = nullptr;
        if (!dev) {
            // HOST_PTRS: ONE persistent launch; the per-instance inputs are staged chunk by chunk on the copy stream while it runs
            // (same protocol as the dense entry points: a 4-byte copy after each chunk publishes how many QPs have landed)
            int chunks = c->opt_chunks;
            if (chunks > count) chunks = count;
            CK(c, cudaMemsetAsync(b->ready_dev, 0, sizeof(int), stream));
            CK(c, cudaEventRecord(b->stage_event, stream));
            CK(c, cudaStreamWaitEvent(c->copy_stream, b->stage_event, 0));
            int bound[96];
            chunks = chunk_bounds(count, chunks, c->prop.multiProcessorCount / 4, bound, 96);  // a wave here: one QP per 4-CTA cluster
            for (int k = 0; k < chunks; ++k) {
                const size_t lo = (size_t)bound[k], hi = (size_t)bound[k + 1];
                CK(c, copy_instances(c->copy_stream, lo, hi - lo));
                b->ready_host[k] = (int)hi;
                CK(c, cudaMemcpyAsync(b->ready_dev, b->ready_host + k, sizeof(int), cudaMemcpyHostToDevice, c->copy_stream));
            }
            ready = b->ready_dev;
        }
        rc = launch_range(b, s, mode, 0, count, dP, dq, nullptr, dl, du, stream, ready, &sp);
        if (rc && !dev) cudaStreamSynchronize(c->copy_stream);
    }
    if (rc) return rc;
    if (!dev) CK(c, cudaStreamSynchronize(stream));
    return SQPB200_OK;
}

extern "C" {

int sqpb200_qp_batch_setup_solve_sparse(sqpb200_qp_batch *b, const sqpb200_qp_settings *s, int count, const double *P,
                                        const double *q, const double *A_values, const int *A_outer, const int *A_inner, int nnz,
                                        int layout, const double *l, const double *u, unsigned flags, void *stream) {
    if (b) {
        b->fact_valid = false;
        b->fused_used = true;
    }
    return run_sparse(b, s, MODE_RESET | MODE_FACTOR | MODE_SOLVE, count, P, q, A_values, A_outer, A_inner, nnz, layout, l, u, flags, stream);
}
int sqpb200_qp_batch_setup_sparse(sqpb200_qp_batch *b, const sqpb200_qp_settings *s, int count, const double *P, const double *q,
                                  const double *A_values, const int *A_outer, const int *A_inner, int nnz, int layout, const double *l,
                                  const double *u, unsigned flags, void *stream) {
    int rc = run_sparse(b, s, MODE_RESET | MODE_FACTOR | MODE_STORE_FACTOR, count, P, q, A_values, A_outer, A_inner, nnz, layout, l, u, flags, stream);
    if (!rc && b) b->fact_valid = true;
    return rc;
}
int sqpb200_qp_batch_update_qp_sparse(sqpb200_qp_batch *b, const sqpb200_qp_settings *s, int count, const double *P, const double *q,
                                      const double *A_values, const int *A_outer, const int *A_inner, int nnz, int layout,
                                      const double *l, const double *u, unsigned flags, void *stream) {
    int rc = run_sparse(b, s, MODE_FACTOR | MODE_STORE_FACTOR, count, P, q, A_values, A_outer, A_inner, nnz, layout, l, u, flags, stream);
    if (!rc && b) b->fact_valid = true;
    return rc;
}
int sqpb200_qp_batch_solve_sparse(sqpb200_qp_batch *b, const sqpb200_qp_settings *s, int count, const double *P, const double *q,
                                  const double *A_values, const int *A_outer, const int *A_inner, int nnz, int layout, const double *l,
                                  const double *u, unsigned flags, void *stream) {
    if (b && !b->fact_valid && b->fused_used)
        return fail(b->ctx, SQPB200_ERR_INVALID,
                    "sqpb200_qp_batch_solve_sparse: the last setup was the fused setup_solve, which does not keep the factor; call setup first");
    return run_sparse(b, s, MODE_LOAD_FACTOR | MODE_SOLVE | MODE_STORE_FACTOR, count, P, q, A_values, A_outer, A_inner, nnz, layout, l, u, flags, stream);
}

int sqpb200_qp_batch_get(sqpb200_qp_batch *b, int count, double *x, double *y, double *z, int *status, int *iter,
                         int *rho_updates, double *rho_estimate, double *res_prim, double *res_dual, unsigned flags,
                         void *stream_) {
    if (!b) return SQPB200_ERR_INVALID;
    sqpb200_ctx *c = b->ctx;
    if (count < 0 || count > b->batch) return fail(c, SQPB200_ERR_INVALID, "count outside [0, batch]");
    CK(c, cudaSetDevice(c->device));
    cudaStream_t stream = (cudaStream_t)stream_;  // NULL is the CUDA legacy default stream
    const cudaMemcpyKind kind = (flags & SQPB200_DEVICE_PTRS) ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
    if (int rc = order_after_last_launch(b, stream)) return rc;
    size_t B = count, n = b->n, m = b->m;
    if (x) CK(c, cudaMemcpyAsync(x, b->x, B * n * sizeof(double), kind, stream));
    if (y && m) CK(c, cudaMemcpyAsync(y, b->y, B * m * sizeof(double), kind, stream));
    if (z && m) CK(c, cudaMemcpyAsync(z, b->z, B * m * sizeof(double), kind, stream));
    if (status) CK(c, cudaMemcpyAsync(status, b->status, B * sizeof(int), kind, stream));
    if (iter) CK(c, cudaMemcpyAsync(iter, b->iter, B * sizeof(int), kind, stream));
    if (rho_updates) CK(c, cudaMemcpyAsync(rho_updates, b->rho_updates, B * sizeof(int), kind, stream));
    if (rho_estimate) CK(c, cudaMemcpyAsync(rho_estimate, b->rho_estimate, B * sizeof(double), kind, stream));
    if (res_prim) CK(c, cudaMemcpyAsync(res_prim, b->res_prim, B * sizeof(double), kind, stream));
    if (res_dual) CK(c, cudaMemcpyAsync(res_dual, b->res_dual, B * sizeof(double), kind, stream));
    if (!(flags & (SQPB200_DEVICE_PTRS | SQPB200_HOST_ASYNC))) CK(c, cudaStreamSynchronize(stream));
    return SQPB200_OK;
}

int sqpb200_qp_batch_set_iterates(sqpb200_qp_batch *b, int count, const double *x, const double *y, const double *z,
                                  unsigned flags, void *stream_) {
    if (!b) return SQPB200_ERR_INVALID;
    sqpb200_ctx *c = b->ctx;
    if (count < 0 || count > b->batch) return fail(c, SQPB200_ERR_INVALID, "count outside [0, batch]");
    CK(c, cudaSetDevice(c->device));
    cudaStream_t stream = (cudaStream_t)stream_;  // NULL is the CUDA legacy default stream
    const cudaMemcpyKind kind = (flags & SQPB200_DEVICE_PTRS) ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
    if (int rc = order_after_last_launch(b, stream)) return rc;
    size_t B = count, n = b->n, m = b->m;
    if (x) CK(c, cudaMemcpyAsync(b->x, x, B * n * sizeof(double), kind, stream));
    if (y && m) CK(c, cudaMemcpyAsync(b->y, y, B * m * sizeof(double), kind, stream));
    if (z && m) CK(c, cudaMemcpyAsync(b->z, z, B * m * sizeof(double), kind, stream));
    if (!(flags & SQPB200_DEVICE_PTRS)) CK(c, cudaStreamSynchronize(stream));
    return SQPB200_OK;
}

int sqpb200_qp_batch_device_view(sqpb200_qp_batch *b, sqpb200_qp_device_view *v) {
    if (!b || !v) return SQPB200_ERR_INVALID;
    v->x = b->x; v->y = b->y; v->z = b->z;
    v->status = b->status; v->iter = b->iter; v->rho_updates = b->rho_updates;
    v->rho_estimate = b->rho_estimate; v->res_prim = b->res_prim; v->res_dual = b->res_dual; v->rho = b->rho;
    v->total_iters = (long long *)b->total_iters;
    return SQPB200_OK;
}

int sqpb200_qp_batch_total_iters(sqpb200_qp_batch *b, long long *total, void *stream_) {
    if (!b || !total) return SQPB200_ERR_INVALID;
    sqpb200_ctx *c = b->ctx;
    CK(c, cudaSetDevice(c->device));
    cudaStream_t stream = (cudaStream_t)stream_;  // NULL is the CUDA legacy default stream
    unsigned long long v = 0;
    if (int rc = order_after_last_launch(b, stream)) return rc;
    CK(c, cudaMemcpyAsync(&v, b->total_iters, sizeof v, cudaMemcpyDeviceToHost, stream));
    CK(c, cudaStreamSynchronize(stream));
    *total = (long long)v;
    return SQPB200_OK;
}

int sqpb200_qp_solve_batch(sqpb200_ctx *ctx, const sqpb200_qp_settings *settings, int batch, int n, int m, const double *P,
                           const double *q, const double *A, const double *l, const double *u, double *x, double *y,
                           double *z, int *status, int *iter, int *rho_updates, double *rho_estimate, double *res_prim,
                           double *res_dual, unsigned flags, void *stream) {
    sqpb200_qp_batch *b = nullptr;
    int rc = sqpb200_qp_batch_create(ctx, batch, n, m, &b);
    if (rc) return rc;
    rc = sqpb200_qp_batch_setup_solve(b, settings, batch, P, q, A, l, u, flags, stream);
    if (!rc) rc = sqpb200_qp_batch_get(b, batch, x, y, z, status, iter, rho_updates, rho_estimate, res_prim, res_dual, flags, stream);
    if (!rc && (flags & SQPB200_DEVICE_PTRS)) {
        cudaError_t e = cudaStreamSynchronize((cudaStream_t)stream);  // the state is freed next
        if (e != cudaSuccess) rc = fail(ctx, SQPB200_ERR_CUDA, "cudaStreamSynchronize", e);
    }
    sqpb200_qp_batch_destroy(b);
    return rc;
}

}  // extern "C"
