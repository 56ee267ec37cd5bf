// quisk_b200/csrc/waterfall.cu -- the waterfall pixel mapper behind the panadapter (SURVEY.md section 8 (f)4;
// quisk.c:5334-5480: watfall_RgbData, watfall_OnGraphData, watfall_GetPixels), batched over many streams.
//
// The reference keeps, per waterfall, a ring of max_height rows (x_origin + width RGB pixels) as a doubly linked list in
// a Python bytearray: OnGraphData steps the current row BACKWARDS (current = current->prior_row), maps the new line of dB
// values to colour indices  l = (int)((dB - gain + yz) * (y_scale + 10) * 0.10 + 128),  yz = 40.0 + y_zero * 0.69,
// clamped to 0..255, and stores the palette's RGB (zero fill past the data); GetPixels walks from the current row along
// next_row, shifting each row by its own x_origin against the requested one, and in scroll mode draws the first seven
// rows 8, 7, ... 2 times (35 lines) before the rest.  Here the ring is an index into [streams][max_height][width * 3]
// bytes in HBM, one ring position for all streams (they are always fed together), the dB lines come straight from
// quisk_cuda_pan_graph's device output, and the pixel block is written on the device for the caller to copy or display.
// Byte-exact against fixtures produced by calling the reference's own methods (tests/golden/make_golden_waterfall.py).
#include <vector>
#include "qc_common.cuh"
#include "../../include/quisk_cuda.h"

namespace qc {

struct Waterfall {
    int S = 0, width = 0, H = 0, cur = 0;
    unsigned char *d_rows = nullptr;    // [S][H][width * 3]
    int *d_xo = nullptr;                // [H] x_origin of every ring row (the same for all streams)
    unsigned char *d_pal = nullptr;     // [3][256] red, green, blue
    void release()
    {
        if (d_rows) cudaFree(d_rows); if (d_xo) cudaFree(d_xo); if (d_pal) cudaFree(d_pal);
        d_rows = nullptr; d_xo = nullptr; d_pal = nullptr;
    }
};

// one new line per stream into ring row `row`
__global__ void waterfall_line_kernel(const double *db, long db_stride, int n_db, unsigned char *rows, int width, int H, int row,
                                      const unsigned char *pal, double gain, double yz, double ys, int *xo, int x_origin)
{
    const int s = blockIdx.y, i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= width) return;
    if (s == 0 && i == 0) xo[row] = x_origin;
    unsigned char *px = rows + (((size_t)s * H + row) * width + i) * 3;
    if (i < n_db) {
        // (dB - gain + yz) * (y_scale + 10) * 0.10 + 128, every operation rounded on its own as the reference's x86-64 code does
        const double d = db[(size_t)s * db_stride + i];
        const double t = __dadd_rn(__dmul_rn(__dmul_rn(__dadd_rn(__dsub_rn(d, gain), yz), ys), 0.10), 128.0);
        int l = (int)t;                                     // truncation toward zero, like the C cast (|t| is far below 2^31 for any dB a graph holds)
        if (!(t > -2147483648.0 && t < 2147483648.0)) l = t > 0 ? 255 : 0;
        l = l < 0 ? 0 : (l > 255 ? 255 : l);
        px[0] = pal[l]; px[1] = pal[256 + l]; px[2] = pal[512 + l];
    } else {
        px[0] = 0; px[1] = 0; px[2] = 0;
    }
}

// output line `ro` of every stream <- ring row src_row[ro], shifted by (x_origin of that row - x_origin asked for) pixels
__global__ void waterfall_pixels_kernel(const unsigned char *rows, const int *xo, int width, int H, int cur, int lines, int scroll,
                                        int x_origin, unsigned char *out, long out_stride)
{
    const int s = blockIdx.z, ro = blockIdx.y, i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= width || ro >= lines) return;
    // which ring row: scroll mode repeats the newest seven rows 8, 7, ... 2 times (quisk.c:5455-5465)
    int k = ro;                                             // rows after the current one
    if (scroll) {
        if (ro < 35) { int acc = 0; k = 0; for (int j = 8; j > 1; j--) { if (ro < acc + j) break; acc += j; k++; } }
        else k = ro - 35 + 7;
    }
    const int r = (cur + k) % H;
    const int sx = i - (xo[r] - x_origin);
    unsigned char *o = out + (size_t)s * out_stride + ((size_t)ro * width + i) * 3;
    if (sx >= 0 && sx < width) {
        const unsigned char *px = rows + (((size_t)s * H + r) * width + sx) * 3;
        o[0] = px[0]; o[1] = px[1]; o[2] = px[2];
    } else {
        o[0] = 0; o[1] = 0; o[2] = 0;
    }
}

}  // namespace qc

using namespace qc;
struct qcWaterfall { qc::Waterfall w; };

extern "C" {

qcWaterfall *quisk_cuda_waterfall_create(int n_streams, int width, int max_height, const unsigned char *red, const unsigned char *green, const unsigned char *blue)
{
    if (ensure_device() != QC_OK) return nullptr;
    if (n_streams <= 0 || width <= 0 || max_height < 2 || !red || !green || !blue) { set_error("waterfall_create: streams, width > 0, max_height >= 2 and three 256-entry palettes"); return nullptr; }
    qcWaterfall *p = new qcWaterfall();
    Waterfall &w = p->w;
    w.S = n_streams; w.width = width; w.H = max_height; w.cur = 0;
    const size_t bytes = (size_t)n_streams * max_height * width * 3;
    if (cudaMalloc((void **)&w.d_rows, bytes) != cudaSuccess || cudaMalloc((void **)&w.d_xo, (size_t)max_height * sizeof(int)) != cudaSuccess ||
        cudaMalloc((void **)&w.d_pal, 768) != cudaSuccess) { set_error("waterfall_create: out of device memory"); w.release(); delete p; return nullptr; }
    cudaMemset(w.d_rows, 0, bytes);                         // watfall_RgbData: every row zeroed, x_origin 0
    cudaMemset(w.d_xo, 0, (size_t)max_height * sizeof(int));
    unsigned char pal[768];
    memcpy(pal, red, 256); memcpy(pal + 256, green, 256); memcpy(pal + 512, blue, 256);
    cudaMemcpy(w.d_pal, pal, 768, cudaMemcpyHostToDevice);
    return p;
}

void quisk_cuda_waterfall_destroy(qcWaterfall *p) { if (p) { p->w.release(); delete p; } }

int quisk_cuda_waterfall_on_graph_data(qcWaterfall *p, const double *d_db, long db_stride, int n_db, int y_zero, int y_scale, double gain, int x_origin, void *stream)
{
    if (!p || !d_db || n_db < 0) { set_error("waterfall_on_graph_data: bad arguments"); return QC_EINVAL; }
    Waterfall &w = p->w;
    cudaStream_t s = (cudaStream_t)stream;
    w.cur = (w.cur + w.H - 1) % w.H;                        // current_row = current_row->prior_row
    const int n = n_db > w.width ? w.width : n_db;
    const double yz = 40.0 + y_zero * 0.69;
    waterfall_line_kernel<<<dim3((w.width + 255) / 256, w.S), 256, 0, s>>>(d_db, db_stride, n, w.d_rows, w.width, w.H, w.cur, w.d_pal, gain, yz, (double)(y_scale + 10), w.d_xo, x_origin);
    count_launch();
    QC_CUDA_LAUNCH();
    return QC_OK;
}

int quisk_cuda_waterfall_get_pixels(qcWaterfall *p, unsigned char *d_pixels, long stream_stride_bytes, int x_origin, int height, int scroll_mode, void *stream)
{
    if (!p || !d_pixels) { set_error("waterfall_get_pixels: bad arguments"); return QC_EINVAL; }
    Waterfall &w = p->w;
    // the reference writes its 35 repeated lines whatever `height` says (and past the caller's buffer if it is smaller): refused here
    if (scroll_mode && height < 35) { set_error("waterfall_get_pixels: scroll mode draws 35 lines before the first plain one; height %d is too small", height); return QC_EINVAL; }
    if (height <= 0) return QC_OK;
    if (stream_stride_bytes < (long)height * w.width * 3) { set_error("waterfall_get_pixels: stream stride %ld is less than height * width * 3", stream_stride_bytes); return QC_EINVAL; }
    waterfall_pixels_kernel<<<dim3((w.width + 127) / 128, height, w.S), 128, 0, (cudaStream_t)stream>>>(w.d_rows, w.d_xo, w.width, w.H, w.cur, height, scroll_mode ? 1 : 0,
                                                                                                         x_origin, d_pixels, stream_stride_bytes);
    count_launch();
    QC_CUDA_LAUNCH();
    return QC_OK;
}

}  // extern "C"
