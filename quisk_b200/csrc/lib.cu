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

const char *quisk_cuda_version(void) { return "libquisk_cuda 0.1 (sm_100a) for Quisk 4.2.52 / WDSP 1.25"; }

}  // extern "C"
