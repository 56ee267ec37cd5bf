// quisk_b200/csrc/ingest.cu -- wire-format ingest: received bytes -> complex double.
//
// The receive path starts with integer samples off the wire; the reference widens them to complex double on the
// host before anything else (quisk.c:2922-2953 for the generic sample source, quisk.c:3746-3763 for Hermes / Metis
// protocol 1).  Doing that step on the device means the host -> device copy carries 2..8 bytes per sample instead of
// 16, which is what bounds the end-to-end rate of the receive chain (PCIe, not the kernels).
//   unpack_iq_kernel     : add_rx_samples.  (I, Q) pairs, 1..4 bytes per component, little or big endian, each
//                          left-justified into an int32 (so every width spans +-2^31), then int -> float -> double
//                          as the reference's `ii + qq * I` does (C's I is a float complex): 4-byte samples keep 24 bits.
//   unpack_hermes_kernel : read_rx_udp10.  1032-byte packets = 8 header bytes + two 512-byte frames (3 sync, 5
//                          control, then records of (1 + multirx) x [3-byte I, 3-byte Q] + 2 microphone bytes);
//                          24-bit big-endian, left-justified; the FIRST triple is the imaginary part (quisk.c:3747-3749).
// Both are bit-exact against the compiled reference loops.
#include "rxchain.h"

namespace qc {

__global__ void unpack_iq_kernel(const unsigned char *__restrict__ in, long byte_stride, int count, int nb, int big,
                                 cd *__restrict__ out, long out_stride)
{
    const int c = blockIdx.y;
    const unsigned char *src = in + (size_t)c * byte_stride;
    cd *dst = out + (size_t)c * out_stride;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += gridDim.x * blockDim.x) {
        const unsigned char *p = src + (size_t)i * 2 * nb;
        unsigned int ii = 0, qq = 0;
        if (nb == 2 && !big && (((size_t)p) & 3) == 0) {                    // the common case: int16 LE pairs
            const unsigned int w = *reinterpret_cast<const unsigned int *>(p);
            ii = w << 16; qq = w & 0xffff0000u;
        } else if (nb == 4 && !big && (((size_t)p) & 7) == 0) {
            const uint2 w = *reinterpret_cast<const uint2 *>(p);
            ii = w.x; qq = w.y;
        } else {
            for (int k = 0; k < nb; k++) {
                // little endian: byte k is bits 8 (4 - nb + k); big endian: byte k is bits 8 (3 - k)
                const int sh = big ? 8 * (3 - k) : 8 * (4 - nb + k);
                ii |= (unsigned int)p[k] << sh;
                qq |= (unsigned int)p[nb + k] << sh;
            }
        }
        // `ii + qq * I` with int operands is FLOAT complex arithmetic in C (quisk.c:2935): int -> float -> double
        dst[i] = make_double2((double)__int2float_rn((int)ii), (double)__int2float_rn((int)qq));
    }
}

__global__ void unpack_hermes_kernel(const unsigned char *__restrict__ pk, int n_packets, int n_rx, int nrec,
                                     cd *__restrict__ out, long out_stride)
{
    const int per_packet = 2 * nrec;
    const long total = (long)n_packets * per_packet;
    const int rec_bytes = n_rx * 6 + 2;
    for (long s = blockIdx.x * (long)blockDim.x + threadIdx.x; s < total; s += (long)gridDim.x * blockDim.x) {
        const long p = s / per_packet;
        const int w = (int)(s - p * per_packet), f = w / nrec, i = w - f * nrec;
        const unsigned char *b = pk + p * 1032 + 11 + 512 * f + 5 + (size_t)i * rec_bytes;
        for (int r = 0; r < n_rx; r++, b += 6) {
            const int xi = (int)((unsigned int)b[0] << 24 | (unsigned int)b[1] << 16 | (unsigned int)b[2] << 8);
            const int xr = (int)((unsigned int)b[3] << 24 | (unsigned int)b[4] << 16 | (unsigned int)b[5] << 8);
            out[(size_t)r * out_stride + s] = make_double2((double)xr, (double)xi);
        }
    }
}

int launch_unpack_iq(const void *d_bytes, long byte_stride, int C, int count, int nb, int big, cd *out, long out_stride, cudaStream_t s)
{
    if (nb < 1 || nb > 4) { set_error("unpack_iq: bytes per component must be 1..4 (got %d)", nb); return QC_EINVAL; }
    if (C <= 0 || count < 0 || byte_stride < (long)count * 2 * nb || out_stride < count) { set_error("unpack_iq: bad sizes"); return QC_EINVAL; }
    if (count == 0) return QC_OK;
    int gx = (count + 255) / 256; if (gx > 4096) gx = 4096;
    unpack_iq_kernel<<<dim3(gx, C), 256, 0, s>>>((const unsigned char *)d_bytes, byte_stride, count, nb, big ? 1 : 0, out, out_stride);
    count_launch();
    QC_CUDA_LAUNCH();
    return QC_OK;
}

// host bytes -> device -> unpack -> chain -> host audio
}  // namespace qc

extern "C" {

int quisk_cuda_unpack_iq(const void *d_bytes, long byte_stride, int n_channels, int count, int bytes, int big_endian,
                         void *d_out, long out_stride, void *stream)
{
    if (qc::ensure_device() != QC_OK) return QC_ENODEV;
    if (!d_bytes || !d_out) { qc::set_error("unpack_iq: null pointer"); return QC_EINVAL; }
    return qc::launch_unpack_iq(d_bytes, byte_stride, n_channels, count, bytes, big_endian, (cd *)d_out, out_stride, (cudaStream_t)stream);
}

int quisk_cuda_hermes_samples_per_packet(int n_rx) { return n_rx >= 1 && n_rx <= 10 ? 2 * (504 / (n_rx * 6 + 2)) : QC_EINVAL; }

int quisk_cuda_unpack_hermes(const void *d_packets, int n_packets, int n_rx, void *d_out, long out_stride, int *n_samples, void *stream)
{
    if (qc::ensure_device() != QC_OK) return QC_ENODEV;
    if (n_rx < 1 || n_rx > 10) { qc::set_error("unpack_hermes: receivers must be 1..10 (got %d)", n_rx); return QC_EINVAL; }
    const int nrec = 504 / (n_rx * 6 + 2);              // quisk.c:3545
    const long total = (long)n_packets * 2 * nrec;
    if (n_samples) *n_samples = (int)total;
    if (n_packets < 0 || !d_packets || !d_out || out_stride < total) { qc::set_error("unpack_hermes: bad arguments"); return QC_EINVAL; }
    if (total == 0) return QC_OK;
    long g = (total + 255) / 256; if (g > 8192) g = 8192;
    qc::unpack_hermes_kernel<<<(int)g, 256, 0, (cudaStream_t)stream>>>((const unsigned char *)d_packets, n_packets, n_rx, nrec, (cd *)d_out, out_stride);
    qc::count_launch();
    QC_CUDA_LAUNCH();
    return QC_OK;
}

}  // extern "C"
