// quisk_b200/csrc/lib.cu -- library status, error reporting, device bring-up.
#include "qc_common.cuh"
#include <cstdarg>

namespace qc {

static thread_local char t_err[512] = "";
std::atomic<unsigned long long> g_launches{0};

void set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(t_err, sizeof(t_err), fmt, ap);
    va_end(ap);
}

int check(cudaError_t e, const char *what, const char *file, int line)
{
    if (e == cudaSuccess) return QC_OK;
    set_error("%s failed: %s (%s:%d)", what, cudaGetErrorString(e), file, line);
    return QC_ECUDA;
}

int ensure_device()
{
    static std::once_flag once;
    static int status = QC_ENODEV;
    std::call_once(once, [] {
        int n = 0;
        cudaError_t e = cudaGetDeviceCount(&n);
        if (e != cudaSuccess || n <= 0) {
            set_error("no usable CUDA device: %s", e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
            status = QC_ENODEV;
            return;
        }
        status = QC_OK;
    });
    if (status != QC_OK && t_err[0] == 0)
        set_error("no usable CUDA device");
    return status;
}

void die_no_device(const char *fn)
{
    fprintf(stderr, "libquisk_cuda: %s: %s -- this library has no CPU fallback\n", fn, t_err[0] ? t_err : "CUDA unavailable");
    abort();
}

}  // namespace qc

extern "C" {

const char *quisk_cuda_last_error(void) { return qc::t_err; }

int quisk_cuda_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int quisk_cuda_set_device(int device)
{
    if (qc::ensure_device() != QC_OK) return QC_ENODEV;
    QC_CUDA(cudaSetDevice(device));
    return QC_OK;
}

unsigned long long quisk_cuda_launch_count(void) { return qc::g_launches.load(); }

// FP64 pipe peak, measured: every thread runs 8 independent DFMA chains, enough warps resident to fill the pipe.
// The roofline SURVEY.md section 8(d) asks to quote next to HBM for the kernels that sit near the FP64 ridge.
static __global__ void __launch_bounds__(256) fp64_peak_kernel(double *out, int iters, double a, double b)
{
    double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; i++) {
        x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
        x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
    }
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
}

int quisk_cuda_fp64_peak(double *dfma_per_second)
{
    if (qc::ensure_device() != QC_OK) return QC_ENODEV;
    if (!dfma_per_second) return QC_EINVAL;
    int dev = 0, n_sm = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
    const int blocks = n_sm * 8, iters = 1 << 15;
    double *d = nullptr;
    QC_CUDA(cudaMalloc((void **)&d, (size_t)blocks * 256 * sizeof(double)));
    cudaEvent_t e0, e1;
    QC_CUDA(cudaEventCreate(&e0)); QC_CUDA(cudaEventCreate(&e1));
    fp64_peak_kernel<<<blocks, 256>>>(d, 1024, 0.999999, 1.0e-9);          // warm-up
    QC_CUDA(cudaEventRecord(e0));
    fp64_peak_kernel<<<blocks, 256>>>(d, iters, 0.999999, 1.0e-9);
    QC_CUDA(cudaEventRecord(e1));
    QC_CUDA(cudaEventSynchronize(e1));
    float ms = 0.f;
    QC_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    *dfma_per_second = (double)blocks * 256.0 * 8.0 * iters / (ms * 1e-3);
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(d);
    qc::count_launch(); qc::count_launch();
    return QC_OK;
}

const char *quisk_cuda_version(void) { return "libquisk_cuda 0.1 (sm_100a) for Quisk 4.2.52 / WDSP 1.25"; }

}  // extern "C"
