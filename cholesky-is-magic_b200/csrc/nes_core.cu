// Context lifetime, accessors, memory ledger and stage timers of libnes.so.
// Mirrors wrapper.c:8-52 (cholmod_allocate/release + 19 get/set accessor pairs) and the
// with-cholmod protocol of sparse-cholesky.lisp:389-406 (allocate -> start -> defaults ... finish
// -> release).
#include <new>

#include <cstdlib>

#include "nes_internal.h"

namespace nes {

int fail(nes_ctx* c, int status, const char* fmt, ...) {
    if (c) {
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(c->err, sizeof(c->err), fmt, ap);
        va_end(ap);
        c->status = status;
    }
    return status;
}

static void ledger_add(nes_ctx* c, void* p, size_t bytes) {
    c->ledger[p] = bytes;
    c->malloc_count += 1;
    c->memory_inuse += bytes;
    if (c->memory_inuse > c->memory_usage) c->memory_usage = c->memory_inuse;
}

static size_t ledger_remove(nes_ctx* c, void* p) {
    auto it = c->ledger.find(p);
    if (it == c->ledger.end()) return 0;
    size_t b = it->second;
    c->ledger.erase(it);
    c->malloc_count -= 1;
    c->memory_inuse -= b;
    return b;
}

void* dev_alloc(nes_ctx* c, size_t bytes) {
    if (!c || !c->started) return nullptr;
    if (bytes == 0) bytes = 16;
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e != cudaSuccess) {
        cudaGetLastError();
        fail(c, NES_ERR_OUT_OF_MEMORY, "cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
        return nullptr;
    }
    ledger_add(c, p, bytes);
    return p;
}

void dev_free(nes_ctx* c, void* p) {
    if (!p) return;
    ledger_remove(c, p);
    cudaFree(p);
}

void* pinned_alloc(nes_ctx* c, size_t bytes) {
    if (!c || !c->started) return nullptr;
    void* p = nullptr;
    cudaError_t e = cudaMallocHost(&p, bytes ? bytes : 16);
    if (e != cudaSuccess) {
        cudaGetLastError();
        fail(c, NES_ERR_OUT_OF_MEMORY, "cudaMallocHost(%zu) failed: %s", bytes,
             cudaGetErrorString(e));
        return nullptr;
    }
    ledger_add(c, p, bytes);
    return p;
}

void pinned_free(nes_ctx* c, void* p) {
    if (!p) return;
    ledger_remove(c, p);
    cudaFreeHost(p);
}

double* ensure_ws(nes_ctx* c, int slot, size_t bytes) {
    if (c->ws_bytes[slot] >= bytes && c->d_ws[slot]) return c->d_ws[slot];
    if (c->d_ws[slot]) {
        cudaStreamSynchronize(c->stream);
        dev_free(c, c->d_ws[slot]);
        c->d_ws[slot] = nullptr;
        c->ws_bytes[slot] = 0;
    }
    size_t want = bytes < (1u << 16) ? (1u << 16) : bytes;
    c->d_ws[slot] = static_cast<double*>(dev_alloc(c, want));
    if (!c->d_ws[slot]) return nullptr;
    c->ws_bytes[slot] = want;
    return c->d_ws[slot];
}

int ensure_pinned(nes_ctx* c, size_t bytes) {
    if (c->pinned_bytes >= bytes) return 0;
    if (c->h_pinned) {
        cudaStreamSynchronize(c->stream);
        pinned_free(c, c->h_pinned);
        c->h_pinned = nullptr;
        c->pinned_bytes = 0;
    }
    size_t want = bytes < 4096 ? 4096 : bytes;
    c->h_pinned = static_cast<double*>(pinned_alloc(c, want));
    if (!c->h_pinned) return c->status;
    c->pinned_bytes = want;
    return 0;
}

int upload(nes_ctx* c, void* dst_dev, const void* src_host, size_t bytes) {
    if (bytes == 0) return 0;
    NES_CUDA(c, cudaMemcpyAsync(dst_dev, src_host, bytes, cudaMemcpyHostToDevice, c->stream));
    // host buffers belong to the caller (Lisp arrays are never pinned, sparse-cholesky.lisp:357):
    // the copy must be complete before we return.
    NES_CUDA(c, cudaStreamSynchronize(c->stream));
    return 0;
}

int download(nes_ctx* c, void* dst_host, const void* src_dev, size_t bytes) {
    if (bytes == 0) return 0;
    NES_CUDA(c, cudaMemcpyAsync(dst_host, src_dev, bytes, cudaMemcpyDeviceToHost, c->stream));
    NES_CUDA(c, cudaStreamSynchronize(c->stream));
    return 0;
}

StageTimer::StageTimer(nes_ctx* ctx, int stage) : c(ctx), idx(-1) {
    if (!c || !c->timing) return;
    auto get_event = [&]() {
        cudaEvent_t e;
        if (!c->event_pool.empty()) {
            e = c->event_pool.back();
            c->event_pool.pop_back();
        } else {
            cudaEventCreate(&e);
        }
        return e;
    };
    nes_ctx::Interval iv;
    iv.stage = stage;
    iv.a = get_event();
    iv.b = get_event();
    cudaEventRecord(iv.a, c->stream);
    c->intervals.push_back(iv);
    idx = static_cast<int>(c->intervals.size()) - 1;
}

StageTimer::~StageTimer() {
    if (idx >= 0) cudaEventRecord(c->intervals[idx].b, c->stream);
}

static void collect_timing(nes_ctx* c) {
    if (c->intervals.empty()) return;
    cudaStreamSynchronize(c->stream);
    for (auto& iv : c->intervals) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, iv.a, iv.b) == cudaSuccess) {
            c->stage_ms[iv.stage] += ms;
            c->stage_count[iv.stage] += 1;
        } else {
            cudaGetLastError();
        }
        c->event_pool.push_back(iv.a);
        c->event_pool.push_back(iv.b);
    }
    c->intervals.clear();
}

}  // namespace nes

using namespace nes;

extern "C" {

nes_ctx* nes_allocate(void) { return new (std::nothrow) nes_ctx(); }

void nes_release(nes_ctx* c) {
    if (!c) return;
    if (c->started) nes_finish(c);
    delete c;
}

int nes_set_device(nes_ctx* c, int device) {
    if (!c) return NES_ERR_INVALID;
    if (c->started) return fail(c, NES_ERR_INVALID, "nes_set_device after nes_start");
    c->device = device;
    return 0;
}

int nes_start(nes_ctx* c) {
    if (!c) return NES_ERR_INVALID;
    if (c->started) return 1;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        cudaGetLastError();
        fail(c, NES_ERR_NO_DEVICE, "no CUDA device: %s (libnes has no CPU fallback)",
             e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
        return 0;
    }
    if (c->device < 0) {
        if (cudaGetDevice(&c->device) != cudaSuccess) c->device = 0;
    }
    if (cudaSetDevice(c->device) != cudaSuccess) {
        fail(c, NES_ERR_NO_DEVICE, "cudaSetDevice(%d) failed", c->device);
        return 0;
    }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, c->device) != cudaSuccess) {
        fail(c, NES_ERR_NO_DEVICE, "cudaGetDeviceProperties failed");
        return 0;
    }
    if (prop.major != 10) {
        fail(c, NES_ERR_NO_DEVICE, "device %d is sm_%d%d; libnes is built for sm_100a only",
             c->device, prop.major, prop.minor);
        return 0;
    }
    c->num_sms = prop.multiProcessorCount;
    int prio_lo = 0, prio_hi = 0;
    cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);  // lo = least urgent (numerically greatest)
    if (getenv("NES_NO_PRIORITY")) prio_lo = prio_hi = 0;    // debugging: every stream at the default priority
    if (cudaStreamCreateWithPriority(&c->stream, cudaStreamNonBlocking, prio_hi) != cudaSuccess ||
        cudaStreamCreateWithPriority(&c->stream_aux, cudaStreamNonBlocking, prio_lo) != cudaSuccess ||
        cudaStreamCreateWithPriority(&c->stream_b, cudaStreamNonBlocking, prio_hi) != cudaSuccess ||
        cudaStreamCreateWithPriority(&c->stream_c, cudaStreamNonBlocking, (prio_lo + prio_hi) / 2) != cudaSuccess ||
        cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&c->ev_panel, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&c->ev_update, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&c->ev_aux, cudaEventDisableTiming) != cudaSuccess) {
        fail(c, NES_ERR_CUDA, "cudaStreamCreate failed");
        return 0;
    }
    c->started = 1;
    c->status = 0;
    return 1;  // CHOLMOD convention: TRUE on success
}

int nes_defaults(nes_ctx* c) {
    if (!c) return 0;
    c->dbound = 0.0;
    c->supernodal_switch = 40.0;
    c->supernodal = 1;
    c->print = 3;
    c->itype = 0;
    c->dtype = 0;
    return 1;
}

int nes_free_work(nes_ctx* c) {
    if (!c || !c->started) return 1;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    for (int i = 0; i < nes_ctx::kNumWs; ++i) {
        if (c->d_ws[i]) dev_free(c, c->d_ws[i]);
        c->d_ws[i] = nullptr;
        c->ws_bytes[i] = 0;
    }
    if (c->d_split_counters) dev_free(c, c->d_split_counters);
    c->d_split_counters = nullptr;
    if (c->h_pinned) pinned_free(c, c->h_pinned);
    c->h_pinned = nullptr;
    c->pinned_bytes = 0;
    return 1;
}

int nes_finish(nes_ctx* c) {
    if (!c) return 0;
    if (!c->started) return 1;
    nes_comm_finalize(c);
    nes_free_work(c);
    collect_timing(c);
    for (auto e : c->event_pool) cudaEventDestroy(e);
    c->event_pool.clear();
    if (c->mark_a) cudaEventDestroy(c->mark_a);
    if (c->mark_b) cudaEventDestroy(c->mark_b);
    c->mark_a = c->mark_b = nullptr;
    if (c->ev_panel) cudaEventDestroy(c->ev_panel);
    if (c->ev_update) cudaEventDestroy(c->ev_update);
    if (c->ev_aux) cudaEventDestroy(c->ev_aux);
    c->ev_panel = c->ev_update = c->ev_aux = nullptr;
    if (c->ev_fork) cudaEventDestroy(c->ev_fork);
    if (c->ev_join) cudaEventDestroy(c->ev_join);
    c->ev_fork = c->ev_join = nullptr;
    if (c->stream_b) cudaStreamDestroy(c->stream_b);
    c->stream_b = nullptr;
    if (c->stream_c) cudaStreamDestroy(c->stream_c);
    c->stream_c = nullptr;
    if (c->stream_aux) cudaStreamDestroy(c->stream_aux);
    c->stream_aux = nullptr;
    if (c->stream) cudaStreamDestroy(c->stream);
    c->stream = nullptr;
    c->started = 0;
    return 1;
}

const char* nes_last_error(const nes_ctx* c) { return c ? c->err : "null context"; }

int nes_version(int version[3]) {
    if (version) {
        version[0] = 0;
        version[1] = 1;
        version[2] = 0;
    }
    return (0 << 16) | (1 << 8) | 0;
}

#define NES_DEFINE_ACCESSOR(FIELD, TYPE)                               \
    TYPE nes_get_##FIELD(const nes_ctx* c) { return (TYPE)(c->FIELD); } \
    TYPE nes_set_##FIELD(nes_ctx* c, TYPE new_value) {                  \
        TYPE old = (TYPE)(c->FIELD);                                    \
        c->FIELD = new_value;                                           \
        return old;                                                     \
    }
NES_DEFINE_ACCESSOR(print, int)
NES_DEFINE_ACCESSOR(print_function, void*)
NES_DEFINE_ACCESSOR(dbound, double)
NES_DEFINE_ACCESSOR(supernodal_switch, double)
NES_DEFINE_ACCESSOR(supernodal, int)
NES_DEFINE_ACCESSOR(selected, int)
NES_DEFINE_ACCESSOR(itype, int)
NES_DEFINE_ACCESSOR(dtype, int)
NES_DEFINE_ACCESSOR(status, int)
NES_DEFINE_ACCESSOR(fl, double)
NES_DEFINE_ACCESSOR(lnz, double)
NES_DEFINE_ACCESSOR(anz, double)
NES_DEFINE_ACCESSOR(modfl, double)
NES_DEFINE_ACCESSOR(malloc_count, size_t)
NES_DEFINE_ACCESSOR(memory_usage, size_t)
NES_DEFINE_ACCESSOR(memory_inuse, size_t)
NES_DEFINE_ACCESSOR(rowfacfl, double)
NES_DEFINE_ACCESSOR(aatfl, double)
NES_DEFINE_ACCESSOR(blas_ok, int)

int nes_get_minor(const nes_ctx* c) { return c ? c->minor : -1; }

int nes_timing_enable(nes_ctx* c, int on) {
    if (!c) return NES_ERR_INVALID;
    c->timing = on ? 1 : 0;
    return 0;
}

int nes_timing_reset(nes_ctx* c) {
    if (!c) return NES_ERR_INVALID;
    if (c->started) collect_timing(c);
    for (int i = 0; i < NES_NUM_STAGES; ++i) {
        c->stage_ms[i] = 0;
        c->stage_count[i] = 0;
    }
    return 0;
}

int nes_timing_get(nes_ctx* c, int stage, double* ms, long long* count) {
    if (!c || stage < 0 || stage >= NES_NUM_STAGES) return NES_ERR_INVALID;
    if (c->started) collect_timing(c);
    if (ms) *ms = c->stage_ms[stage];
    if (count) *count = c->stage_count[stage];
    return 0;
}

long long nes_get_launch_count(const nes_ctx* c) { return c ? c->launches : 0; }
double nes_get_form_flops(const nes_ctx* c) { return c ? c->form_flops : 0.0; }
int nes_set_ordering_leaf(nes_ctx* c, int leaf) {
    if (!c) return 0;
    const int old = c->nd_leaf;
    c->nd_leaf = leaf < 0 ? 0 : leaf;
    return old;
}

int nes_mark_begin(nes_ctx* c) {
    NES_ENTER(c);
    if (!c->mark_a) {
        NES_CUDA(c, cudaEventCreate(&c->mark_a));
        NES_CUDA(c, cudaEventCreate(&c->mark_b));
    }
    NES_CUDA(c, cudaStreamSynchronize(c->stream));
    NES_CUDA(c, cudaEventRecord(c->mark_a, c->stream));
    return 0;
}

int nes_mark_end(nes_ctx* c, double* ms) {
    NES_ENTER(c);
    if (!c->mark_a || !ms) return fail(c, NES_ERR_INVALID, "nes_mark_end without nes_mark_begin");
    NES_CUDA(c, cudaEventRecord(c->mark_b, c->stream));
    NES_CUDA(c, cudaEventSynchronize(c->mark_b));
    float f = 0.f;
    NES_CUDA(c, cudaEventElapsedTime(&f, c->mark_a, c->mark_b));
    *ms = f;
    return 0;
}

int nes_synchronize(nes_ctx* c) {
    if (!c || !c->started) return NES_ERR_NO_DEVICE;
    NES_CUDA(c, cudaStreamSynchronize(c->stream));
    return 0;
}

}  // extern "C"
