// projectultra_b200/csrc/pu_ctx.cu — context, error reporting, scratch buffers.
#include "pu_internal.h"

namespace pu {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

pu_status Buffer::reserve(size_t bytes) {
    if (bytes <= cap) return PU_OK;
    release();
    size_t want = bytes + bytes / 4 + 256;
    cudaError_t e = pinned_host ? cudaMallocHost(&ptr, want) : cudaMalloc(&ptr, want);
    if (e != cudaSuccess) {
        ptr = nullptr;
        cap = 0;
        set_error("allocation of %zu bytes failed: %s", want, cudaGetErrorString(e));
        return PU_ERR_NOMEM;
    }
    cap = want;
    return PU_OK;
}

void Buffer::release() {
    if (ptr) {
        if (pinned_host) cudaFreeHost(ptr);
        else cudaFree(ptr);
    }
    ptr = nullptr;
    cap = 0;
}

}  // namespace pu

extern "C" {

int pu_abi_version(void) { return 1; }

const char* pu_status_string(pu_status s) {
    switch (s) {
        case PU_OK: return "ok";
        case PU_ERR_INVALID: return "invalid argument";
        case PU_ERR_CUDA: return "CUDA error";
        case PU_ERR_NOMEM: return "out of memory";
        case PU_ERR_UNSUPPORTED: return "unsupported configuration";
    }
    return "unknown";
}

const char* pu_last_error(void) { return pu::g_err; }

pu_status pu_init(int device, pu_ctx** out) {
    PU_REQUIRE(out != nullptr, "pu_init: out is NULL");
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        pu::set_error("pu_init: no CUDA device (%s); this library has no CPU fallback",
                      e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
        return PU_ERR_CUDA;
    }
    PU_REQUIRE(device >= 0 && device < n, "pu_init: device index out of range");
    PU_CUDA_TRY(cudaSetDevice(device));
    cudaDeviceProp prop;
    PU_CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
        pu::set_error("pu_init: device %d is sm_%d%d; kernels are built for sm_100a only", device, prop.major, prop.minor);
        return PU_ERR_CUDA;
    }
    pu_ctx* c = new pu_ctx();
    c->device = device;
    c->sm_count = prop.multiProcessorCount;
    c->cc_major = prop.major;
    c->cc_minor = prop.minor;
    c->h_in.pinned_host = true;
    c->h_out.pinned_host = true;
    e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) {
        delete c;
        pu::set_error("pu_init: cudaStreamCreate failed: %s", cudaGetErrorString(e));
        return PU_ERR_CUDA;
    }
    for (auto& sl : c->pipe) {
        sl.h_in.pinned_host = true;
        sl.h_out.pinned_host = true;
        if (cudaStreamCreateWithFlags(&sl.stream, cudaStreamNonBlocking) != cudaSuccess) sl.stream = nullptr;
    }
    if (cudaEventCreateWithFlags(&c->pipe_ev, cudaEventDisableTiming) != cudaSuccess) c->pipe_ev = nullptr;
    if (!c->pipe[0].stream || !c->pipe[1].stream || !c->pipe_ev) {
        pu::set_error("pu_init: could not create the pipeline streams");
        pu_destroy(c);
        return PU_ERR_CUDA;
    }
    *out = c;
    return PU_OK;
}

void pu_destroy(pu_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    c->d_in.release();
    c->d_out.release();
    c->d_aux.release();
    c->h_in.release();
    c->h_out.release();
    c->f_llr.release(); c->f_bytes.release(); c->f_ok.release(); c->f_iters.release(); c->f_out.release();
    for (auto& sl : c->pipe) {
        if (sl.stream) cudaStreamSynchronize(sl.stream);
        sl.d_in.release(); sl.d_llr.release(); sl.d_out.release(); sl.h_in.release(); sl.h_out.release();
        if (sl.stream) cudaStreamDestroy(sl.stream);
    }
    if (c->pipe_ev) cudaEventDestroy(c->pipe_ev);
    c->cfo_mix.release();
    for (auto& b : c->sweep) b.release();
    for (auto& e : c->sweep_ev) if (e) cudaEventDestroy(e);
    cudaStreamDestroy(c->stream);
    delete c;
}

int pu_device_sm_count(const pu_ctx* c) { return c ? c->sm_count : 0; }

pu_status pu_synchronize(pu_ctx* c, void* stream) {
    PU_REQUIRE(c != nullptr, "pu_synchronize: ctx is NULL");
    PU_CUDA_TRY(cudaStreamSynchronize(pu::pick_stream(c, stream)));
    return PU_OK;
}

uint64_t pu_kernel_launches(const pu_ctx* c) { return c ? c->launches.load() : 0; }

void pu_transfer_bytes(const pu_ctx* c, uint64_t* h2d, uint64_t* d2h) {
    if (h2d) *h2d = c ? c->h2d_bytes.load() : 0;
    if (d2h) *d2h = c ? c->d2h_bytes.load() : 0;
}

}  // extern "C"
