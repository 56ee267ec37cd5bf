// quisk_b200/csrc/rxfused.cu -- the fused full-rate decimator: ONE kernel for the tuning NCO
// (quisk.c:2477-2488) and every quisk_cDecim2HB45 / quisk_cDecimate stage of
// quisk_process_decimate (+ the demodulator's own pre-decimation, quisk.c:1909-1983).
//
// Design (B200-first, see DESIGN.md section "fused decimator"):
//  * One CTA streams one channel through time.  Per chunk of T0 input samples it loads the
//    chunk with coalesced 16-byte loads (multiplying by the NCO phasor on the way in), then
//    runs every stage out of shared memory into the next stage's shared-memory buffer; only
//    the last stage writes HBM.  Algorithmic traffic is therefore 16 B in + 16/Dtot B out per
//    input sample, and nothing is re-read: there are no halos because a channel never leaves
//    its CTA, and the inter-chunk FIR histories stay in shared memory.
//  * Between calls the per-stage histories live in the same [C][H] arrays the unfused exact
//    kernels use (batch.cu), so a stream can switch between the two paths at any call.
//  * Half-band stages are register blocked: a thread produces R consecutive outputs from one
//    window of 22+R-1 even-phase and R odd-phase samples held in registers (R = 8/4/2), so
//    shared-memory traffic per output drops from 23 to (21+2R)/R loads.  Buffers are padded
//    by one element every 2R so that the per-thread stride (2R+1 elements) is odd and the
//    16-byte loads of a quarter-warp never collide; the buffer origin is shifted so every
//    thread's window starts on a pad boundary and all offsets are compile-time immediates.
//  * FIR stages split taps across TS adjacent lanes (13 taps per step, R = 4 outputs per
//    thread from one register window), then butterfly-reduce the partial sums with shuffles.
//  * Arithmetic is FP64 FMA in (nearly) the reference's summation order: within ~1e-15 of
//    the exact path, not bit-identical to it (tests/test_batch_gpu.py states the bound).
#include "rxchain.h"
#include "nco_device.cuh"

namespace qc {

static constexpr int FNT = 128;         // chunk granularity: CTAs are 128 or 256 threads wide
static constexpr int MAXST = 10;
static constexpr int FIR_KB = 13;       // taps per register window (odd: lanes of a tap split never collide)
static constexpr int FIR_R = 2;
static constexpr int MAXCOEF = 1536;

struct FStage {
    int type;           // 0 = half band, 1 = FIR
    int D, nTaps;
    int Hs;             // history length of the state arrays (BatchFilter::H)
    int Ha;             // history kept in shared memory (>= Hs, covers the zero-padded taps)
    int u0;             // offset of the first output-producing sample among the new samples
    int R;
    int split;          // half band: 1 = a lane PAIR shares R outputs, one lane per component (hb_stage_split)
    int pu;             // pad unit (0 = unpadded)
    unsigned magic;     // ceil(2^32 / pu)
    int org;            // origin shift: q = logical + org
    int buf;            // offset of this stage's input buffer in shared memory (cd units)
    int buf_len;        // physical length (cd units)
    int coef;           // FIR: offset into FusedParams::coef
    int Kpad, TS;       // FIR: zero-padded tap count, tap splits
    int Rplan;          // FIR: outputs per thread in the plan-specialised kernels
    int p0;             // half band: physical index of thread 0's window start
    int n_full;         // new samples this stage sees per full chunk
    int n_out_full;     // outputs per full chunk
    const cd *hin;
    cd *hout;
};

struct FusedParams {
    int ns;
    FStage st[MAXST];
    const cd *in; long in_stride; int n_in;
    cd *out; long out_stride;
    int T0;
    const double *nco; const cd *vstart; unsigned long long n_base; int tune;
    int smem_cd;        // total shared memory in cd units
    int scratch;        // offset of the history-slide scratch area (cd units)
    int coef_sm;        // offset of the tap copy in shared memory (cd units)
    int ncoef;
    int ab_stride;      // three-group kernel: distance (cd units) between the two halves of stage 1's input buffer
    int tw_stride;      // tail-warp kernel: distance (cd units) between the two halves of the first tail stage's input buffer
    int deepk;          // multi-rate plans: the deep stages run once per this many chunks
    long long *trace;   // optional [C][16 chunks][16] clock64() stamps (debug)
    double coef[MAXCOEF];
};

__constant__ double c_hb[12] = {        // filter.c:381-384
    0.000018566625444266, -0.000118469698701817, 0.000457318798253456,
    -0.001347840471412094, 0.003321838571445455, -0.007198422696929033,
    0.014211106939802483, -0.026424776824073383, 0.048414810444971007,
    -0.096214669073304823, 0.314881034738348550, 0.500000000000000000 };

__device__ __forceinline__ int phys(const FStage &s, int logical)
{
    const unsigned q = (unsigned)(logical + s.org);
    return (int)(q + __umulhi(q, s.magic));
}

struct Sink {           // where a stage's outputs go: the next stage's buffer or HBM
    cd *sm;             // shared memory base (nullptr -> global)
    int H, org; unsigned magic;
    cd *g;
    __device__ __forceinline__ void put(int m, cd v) const
    {
        if (sm) {
            const unsigned q = (unsigned)(H + m + org);
            sm[q + __umulhi(q, magic)] = v;
        } else {
            g[m] = v;
        }
    }
    template <bool TOGLOBAL>
    __device__ __forceinline__ void put_c(int m, int comp, double v) const      // one component of output m
    {
        if constexpr (!TOGLOBAL) {
            const unsigned q = (unsigned)(H + m + org);
            double *d = reinterpret_cast<double *>(sm);
            d[2 * (q + __umulhi(q, magic)) + comp] = v;
        } else {
            reinterpret_cast<double *>(g)[2 * m + comp] = v;
        }
    }
};

__device__ __forceinline__ cd fmaz(cd a, double c, cd acc) { return make_double2(fma(a.x, c, acc.x), fma(a.y, c, acc.y)); }

// Half band, R outputs per thread.  sb = stage buffer, p0 = physical index of thread 0's window start.
// The window is streamed: every even-phase sample E[j] = X[n0 - 42 + 2j] is loaded once and fed to
// each of the (up to R) outputs it belongs to, so the thread holds R accumulators and a couple of
// loads in flight instead of the whole window.  samples[k] of output r is E[r + 21 - k] and meets
// coef[min(k, 21-k)] (filter.c:401-413 without the pre-addition of the symmetric pair: same number
// of FP64 operations, a quarter of the registers).
template <int R>
__device__ __forceinline__ void hb_stage(const cd *__restrict__ sb, int p0, int n_out, const Sink &sink, const int t)
{
    const int m0 = t * R;
    if (m0 >= n_out) return;
    const cd *w = sb + p0 + (2 * R + 1) * t;
    cd acc[R];
#pragma unroll
    for (int r = 0; r < R; r++) acc[r] = make_double2(0.0, 0.0);
#pragma unroll
    for (int j = 0; j < 22 + R - 1; j++) {
        const cd e = w[2 * j + (2 * j) / (2 * R)];
#pragma unroll
        for (int r = 0; r < R; r++) {
            const int k = r + 21 - j;                    // samples[] index of E[j] for output r
            if (k >= 0 && k <= 21) acc[r] = fmaz(e, c_hb[k <= 10 ? k : 21 - k], acc[r]);
        }
    }
#pragma unroll
    for (int r = 0; r < R; r++) {
        const cd o = w[(21 + 2 * r) + (21 + 2 * r) / (2 * R)];     // center[10] = X[n - 21]
        acc[r] = fmaz(o, c_hb[11], acc[r]);
        if (m0 + r < n_out) sink.put(m0 + r, acc[r]);
    }
}

// Half band with the two components of a sample on two adjacent lanes: lane 2p handles the real parts and lane 2p+1
// the imaginary parts of outputs R p .. R p + R - 1.  For the same number of accumulator registers a lane covers
// twice as many consecutive outputs as hb_stage, so the 21-sample halo of its window is amortised over twice the
// work: (21 + 2R) 8-byte loads per R component outputs, i.e. 53 x 8 B per 16 at R = 16 against 37 x 16 B per 8
// complex outputs at R = 8 -- 28 % fewer shared-memory wavefronts in stage 0, 36-42 % in the deeper stages.  The
// pair's window starts every 2R+1 elements = 8R+4 words, so a half-warp's 8-byte loads cover all 32 banks once.
// Same FMA order per component as hb_stage: the two variants are bit-identical.
template <int R, bool TOGLOBAL>
__device__ __forceinline__ void hb_stage_split(const cd *__restrict__ sb, int p0, int n_out, const Sink &sink, const int t)
{
    const int pr = t >> 1, comp = t & 1;
    const int m0 = pr * R;
    if (m0 >= n_out) return;
    const double *w = reinterpret_cast<const double *>(sb + p0 + (2 * R + 1) * pr) + comp;
    double acc[R];
#pragma unroll
    for (int r = 0; r < R; r++) acc[r] = 0.0;
#pragma unroll
    for (int j = 0; j < 22 + R - 1; j++) {
        const double e = w[2 * (2 * j + (2 * j) / (2 * R))];
#pragma unroll
        for (int r = 0; r < R; r++) {
            const int k = r + 21 - j;
            if (k >= 0 && k <= 21) acc[r] = fma(e, c_hb[k <= 10 ? k : 21 - k], acc[r]);
        }
    }
#pragma unroll
    for (int r = 0; r < R; r++) {
        const double o = w[2 * ((21 + 2 * r) + (21 + 2 * r) / (2 * R))];
        acc[r] = fma(o, c_hb[11], acc[r]);
        if (m0 + r < n_out) sink.template put_c<TOGLOBAL>(m0 + r, comp, acc[r]);
    }
}

// FIR with decimation D: lane = g * TS + ts; thread handles outputs R*g .. R*g+R-1 over its tap blocks.
template <int D, int NT>
__device__ __forceinline__ void fir_stage(const cd *__restrict__ sb, const FStage &s, const double *__restrict__ coef,
                                          int n_out, const Sink &sink)
{
    constexpr int R = FIR_R, KB = FIR_KB, W = KB + D * (R - 1);
    const int TS = s.TS;
    const int t = threadIdx.x;
    const int ts = t & (TS - 1);
    int g = t / TS;
    const int n_groups = (n_out + R - 1) / R;
    const int rounds = (n_groups * TS + NT - 1) / NT;           // uniform
    for (int round = 0; round < rounds; round++, g += NT / TS) {
        const bool live = g < n_groups;
        const int gg = live ? g : 0;
        cd acc[R];
#pragma unroll
        for (int r = 0; r < R; r++) acc[r] = make_double2(0.0, 0.0);
        // logical index of the sample that meets tap 0 for output r = 0
        const int nbase = s.Hs + s.u0 + D * R * gg + (s.Ha - s.Hs);
        for (int k0 = ts * KB; k0 < s.Kpad; k0 += TS * KB) {
            const cd *w = sb + (nbase - k0 - (KB - 1) + s.org);
            cd Wn[W];
#pragma unroll
            for (int i = 0; i < W; i++) Wn[i] = w[i];
#pragma unroll
            for (int kk = 0; kk < KB; kk++) {
                const double c = coef[k0 + kk];
#pragma unroll
                for (int r = 0; r < R; r++) acc[r] = fmaz(Wn[D * r + KB - 1 - kk], c, acc[r]);
            }
        }
        for (int off = TS >> 1; off > 0; off >>= 1) {
#pragma unroll
            for (int r = 0; r < R; r++) {
                acc[r].x += __shfl_xor_sync(0xffffffffu, acc[r].x, off);
                acc[r].y += __shfl_xor_sync(0xffffffffu, acc[r].y, off);
            }
        }
        if (live && ts == 0) {
#pragma unroll
            for (int r = 0; r < R; r++)
                if (R * g + r < n_out) sink.put(R * g + r, acc[r]);
        }
    }
}

// Plan-specialised FIR: the whole (zero-padded) tap table is TS * KB long, so a lane's KB taps never
// change -- they are loaded into registers once per kernel (cf) -- and one register window of
// KB + D (R - 1) samples feeds R outputs.  lane = g * TS + ts.
template <int D, int NT, int R, bool TOGLOBAL>
__device__ __forceinline__ void fir_stage_c(const cd *__restrict__ sb, const FStage &s, const double (&cf)[FIR_KB],
                                            int n_out, const Sink &sink, const int t)
{
    constexpr int KB = FIR_KB, W = KB + D * (R - 1), TS = 8, V = 2 * R;
    const int ts = t & (TS - 1);
    int g = t / TS;
    const int n_groups = (n_out + R - 1) / R;
    const int rounds = (n_groups * TS + NT - 1) / NT;           // uniform
    for (int round = 0; round < rounds; round++, g += NT / TS) {
        const bool live = g < n_groups;
        const int gg = live ? g : 0;
        cd acc[R];
#pragma unroll
        for (int r = 0; r < R; r++) acc[r] = make_double2(0.0, 0.0);
        const cd *w = sb + (s.Ha + s.u0 + D * R * gg - ts * KB - (KB - 1) + s.org);
        cd Wn[W];
#pragma unroll
        for (int i = 0; i < W; i++) Wn[i] = w[i];
#pragma unroll
        for (int kk = 0; kk < KB; kk++) {
#pragma unroll
            for (int r = 0; r < R; r++) acc[r] = fmaz(Wn[D * r + KB - 1 - kk], cf[kk], acc[r]);
        }
        // Sum over the 8 tap lanes as a reduce-scatter: in the round with lane distance `off` a lane keeps one half of
        // its value list and hands the other half to its partner, so 2R values cost 2R - 1 exchanges (7 for R = 4)
        // instead of the 3 x 2R of an all-reduce butterfly, and lane ts ends up with value ts (output ts / 2,
        // component ts & 1), which it stores itself.  Same pairs, same tree: bit-identical sums.
        double v[V];
#pragma unroll
        for (int r = 0; r < R; r++) { v[2 * r] = acc[r].x; v[2 * r + 1] = acc[r].y; }
        int idx = 0;
        int n = V;
#pragma unroll
        for (int off = TS >> 1; off > 0; off >>= 1) {
            const bool up = (ts & off) != 0;
            if (n >= 2) {
                const int half = n / 2;
#pragma unroll
                for (int i = 0; i < V / 2; i++) {
                    if (i < half) {
                        const double send = up ? v[i] : v[i + half];
                        const double keep = up ? v[i + half] : v[i];
                        v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
                    }
                }
                idx += up ? half : 0;
                n = half;
            } else {
                v[0] += __shfl_xor_sync(0xffffffffu, v[0], off);
            }
        }
        // lanes that share a value after the plain butterfly rounds: the lowest one stores
        constexpr int DUP = V >= TS ? 1 : TS / V;
        if (live && (ts & (DUP - 1)) == 0 && R * g + (idx >> 1) < n_out) sink.template put_c<TOGLOBAL>(R * g + (idx >> 1), idx & 1, v[0]);
    }
}

// register-resident tap sets for the (at most MAXFIR) FIR stages of a plan
static constexpr int MAXFIR = 2;
struct FirTaps { double cf[MAXFIR][FIR_KB]; };

__device__ __forceinline__ int stage_out_count(const FStage &s, int n_in)
{
    return n_in > s.u0 ? (n_in - s.u0 - 1) / s.D + 1 : 0;
}

// ---- plan-specialised cascade: stage kinds are template constants (CODE = type*100 + R*10 + D), the
//      stage index is static, so every descriptor field is a constant-bank operand and the chunk
//      loop is straight-line code.  A partial (ragged) chunk uses the same code with runtime counts.
template <int NT, int CODE, int IDX, bool LAST, int FIRIDX>
__device__ __forceinline__ void run_stage_c(cd *sm, const FusedParams &P, int n_out, cd *gdst, const FirTaps &ft, int sink_off)
{
    constexpr int TYPE = CODE / 100, R = (CODE / 10) % 10, D = CODE % 10;
    const FStage &S = P.st[IDX];
    Sink sink;
    if constexpr (!LAST) {
        const FStage &N = P.st[IDX + 1];
        sink.sm = sm + N.buf; sink.H = N.Ha + sink_off; sink.org = N.org; sink.magic = N.magic; sink.g = nullptr;
    } else {
        sink.sm = nullptr; sink.H = 0; sink.org = 0; sink.magic = 0; sink.g = gdst;
    }
    const cd *sb = sm + S.buf;
    if constexpr (TYPE == 0) {
        hb_stage<R>(sb, S.p0, n_out, sink, threadIdx.x);
    } else if constexpr (TYPE == 2) {           // CODE's R digit is the complex-equivalent blocking: a lane runs 2 R outputs
        hb_stage_split<2 * R, LAST>(sb, S.p0, n_out, sink, threadIdx.x);
    } else {
        fir_stage_c<D, NT, R, LAST>(sb, S, ft.cf[FIRIDX], n_out, sink, threadIdx.x);
    }
}

// ---- the same cascade for the tail-warp kernel: NT threads numbered t = 0 .. NT-1 run stages [START, STOP) and meet at
//      a barrier of their own kind (SYNC 1: named barrier 1 of the 128 main threads; SYNC 2: the tail warp alone).
//      src_off shifts the input buffer of stage START, sink_off the output position of stage STOP-1 (the two halves of
//      the double buffer between the main warps and the tail warp).
template <int SYNC> __device__ __forceinline__ void group_sync()
{
    if constexpr (SYNC == 1) asm volatile("bar.sync 1, 128;" ::: "memory");
    else if constexpr (SYNC == 2) asm volatile("bar.sync 6, 64;" ::: "memory");
    else if constexpr (SYNC == 3) asm volatile("bar.sync 7, 64;" ::: "memory");
    else __syncthreads();
}

template <int NT, int SYNC, bool FULL, int START, int STOP, int IDX, int FIRIDX, int CODE0, int... REST>
__device__ __forceinline__ int cascade_x(cd *sm, const FusedParams &P, int n_in, cd *gdst, const FirTaps &ft, const int t, int src_off, int sink_off)
{
    constexpr int TYPE = CODE0 / 100, R = (CODE0 / 10) % 10, D = CODE0 % 10;
    constexpr int NEXTFIR = FIRIDX + (TYPE == 1 ? 1 : 0);
    constexpr bool LAST = sizeof...(REST) == 0;
    int n_out = n_in;
    if constexpr (IDX >= START && IDX < STOP) {
        const FStage &S = P.st[IDX];
        n_out = FULL ? S.n_out_full : (n_in > S.u0 ? (n_in - S.u0 - 1) / D + 1 : 0);
        Sink sink;
        if constexpr (!LAST) {
            const FStage &N = P.st[IDX + 1];
            sink.sm = sm + N.buf + (IDX == STOP - 1 ? sink_off : 0); sink.H = N.Ha; sink.org = N.org; sink.magic = N.magic; sink.g = nullptr;
        } else {
            sink.sm = nullptr; sink.H = 0; sink.org = 0; sink.magic = 0; sink.g = gdst;
        }
        const cd *sb = sm + S.buf + (IDX == START ? src_off : 0);
        if constexpr (TYPE == 0) hb_stage<R>(sb, S.p0, n_out, sink, t);
        else if constexpr (TYPE == 2) hb_stage_split<2 * R, LAST>(sb, S.p0, n_out, sink, t);
        else fir_stage_c<D, NT, R, LAST>(sb, S, ft.cf[FIRIDX], n_out, sink, t);
        group_sync<SYNC>();
    }
    if constexpr (sizeof...(REST) > 0 && IDX + 1 < STOP)
        return cascade_x<NT, SYNC, FULL, START, STOP, IDX + 1, NEXTFIR, REST...>(sm, P, n_out, gdst, ft, t, src_off, sink_off);
    else
        return n_out;
}

// Runs stages [START, STOP) of the plan.  n_in is the input count of stage START; sink_off is where, behind
// the history of stage STOP (if there is one), the outputs of stage STOP-1 go (the deep stages of a multi-rate
// plan accumulate several chunks before they run).
template <int NT, bool FULL, int START, int STOP, int IDX, int FIRIDX, int CODE0, int... REST>
__device__ __forceinline__ int cascade_c(cd *sm, const FusedParams &P, int n_in, cd *gdst, long long *tr, const FirTaps &ft, int sink_off)
{
    constexpr int D = CODE0 % 10;
    constexpr int NEXTFIR = FIRIDX + (CODE0 / 100 == 1 ? 1 : 0);
    int n_out = n_in;
    if constexpr (IDX >= START && IDX < STOP) {
        const FStage &S = P.st[IDX];
        n_out = FULL ? S.n_out_full : (n_in > S.u0 ? (n_in - S.u0 - 1) / D + 1 : 0);
        run_stage_c<NT, CODE0, IDX, sizeof...(REST) == 0, FIRIDX>(sm, P, n_out, gdst, ft, IDX == STOP - 1 ? sink_off : 0);
        __syncthreads();
        if (FULL && tr) tr[3 + IDX] = clock64();
    }
    if constexpr (sizeof...(REST) > 0 && IDX + 1 < STOP)
        return cascade_c<NT, FULL, START, STOP, IDX + 1, NEXTFIR, REST...>(sm, P, n_out, gdst, tr, ft, sink_off);
    else
        return n_out;
}

// history slide table: element idx (counted over stages [s0, s1)) -> (source, destination) in shared memory
__device__ __forceinline__ void slide_entry(const FusedParams &P, int s0, int s1, int idx, int &src, int &dst)
{
    src = -1; dst = -1;
    for (int s = s0; s < s1; s++) {
        const FStage &S = P.st[s];
        if (idx >= 0 && idx < S.Ha) { src = S.buf + phys(S, S.n_full + idx); dst = S.buf + phys(S, idx); }
        idx -= S.Ha;
    }
}

// One stage of the cascade for n_out outputs (uniform across the CTA).
template <int NT>
__device__ __forceinline__ void run_stage(cd *sm, const FusedParams &P, int s, int n_out, cd *gdst)
{
    const FStage &S = P.st[s];
    Sink sink;
    if (s + 1 < P.ns) {
        const FStage &N = P.st[s + 1];
        sink.sm = sm + N.buf; sink.H = N.Ha; sink.org = N.org; sink.magic = N.magic; sink.g = nullptr;
    } else {
        sink.sm = nullptr; sink.H = 0; sink.org = 0; sink.magic = 0; sink.g = gdst;
    }
    const cd *sb = sm + S.buf;
    if (S.type == 0) {
        if (S.R == 8) { if constexpr (NT == 128) hb_stage<8>(sb, S.p0, n_out, sink, threadIdx.x); }
        else if (S.R == 4) hb_stage<4>(sb, S.p0, n_out, sink, threadIdx.x);
        else hb_stage<2>(sb, S.p0, n_out, sink, threadIdx.x);
    } else {
        const double *coef = reinterpret_cast<const double *>(sm + P.coef_sm) + S.coef;
        if (S.D == 2) fir_stage<2, NT>(sb, S, coef, n_out, sink);
        else if (S.D == 3) fir_stage<3, NT>(sb, S, coef, n_out, sink);
        else if (S.D == 5) fir_stage<5, NT>(sb, S, coef, n_out, sink);
        else fir_stage<1, NT>(sb, S, coef, n_out, sink);
    }
}

// SPLIT: plan kernels only -- stages [SPLIT, ns) are "deep": they run once every P.deepk chunks on the
// accumulated output of stage SPLIT-1, which amortises their fixed per-phase latency (they carry <10 % of
// the flops but cost ~30 % of a chunk's time when run every chunk).  SPLIT == number of stages: no deep part.
template <int... P> struct PlanFirst { static constexpr int code = 0; };
template <int C0, int... P> struct PlanFirst<C0, P...> { static constexpr int code = C0; };

template <int NT, int R0, int MINB, int SPLIT, int... PLAN>
__global__ void __launch_bounds__(NT, MINB) fused_decim_kernel(const __grid_constant__ FusedParams P)
{
    extern __shared__ double smem_raw[];
    cd *sm = reinterpret_cast<cd *>(smem_raw);
    const int c = blockIdx.x;
    const int tid = threadIdx.x;
    __shared__ cd s_pstep;                      // phase^NT
    // loads per thread per chunk: a plan kernel's chunk is what its first stage consumes in one round, 2 R0 NT samples
    // (2048 for <128, 8> and <256, 4>, 1024 for <128, 4>); the generic kernel takes any T0 <= 2048
    constexpr int NLD = sizeof...(PLAN) > 0 ? 2 * R0 : 2048 / NT;
    constexpr int NSL = 512 / NT;               // history elements a thread carries in the slide

    // zero everything once: pads, and the history beyond what the state arrays hold
    for (int i = tid; i < P.smem_cd; i += NT) sm[i] = make_double2(0.0, 0.0);
    __syncthreads();
    for (int s = 0; s < P.ns; s++) {
        const FStage &S = P.st[s];
        const cd *h = S.hin + (size_t)c * S.Hs;
        for (int i = tid; i < S.Hs; i += NT) sm[S.buf + phys(S, (S.Ha - S.Hs) + i)] = h[i];
    }
    {
        double *cs = reinterpret_cast<double *>(sm + P.coef_sm);
        for (int i = tid; i < P.ncoef; i += NT) cs[i] = P.coef[i];
    }
    // NCO: sample k*NT + tid of a chunk is multiplied by u * s_q[k], u = v(chunk start + tid),
    // s_q[k] = phase^(k NT); u advances by pstep = phase^T0 per chunk.  No per-sample recurrence.
    __shared__ cd s_q[2048 / NT];
    cd u = make_double2(1.0, 0.0), pstep = make_double2(1.0, 0.0);
    if (P.tune) {
        const double *nc = P.nco + (size_t)c * 8;
        if (tid == 0) s_pstep = nco_pow(nc, (unsigned long long)P.T0);
        if (tid < 2048 / NT) s_q[tid] = nco_pow(nc, (unsigned long long)tid * NT);
        u = cmul_rn(P.vstart[c], nco_pow(nc, P.n_base + (unsigned long long)tid));
    }
    // slide table for full chunks: which shared-memory element this thread moves where
    constexpr int NS = (int)sizeof...(PLAN);
    constexpr bool MULTI = NS > 0 && SPLIT < NS;
    constexpr int NSL_D = MULTI ? 3 : 1;        // deep-stage slide slots (sum of their Ha <= 3 * NT)
    int sl_src[NSL], sl_dst[NSL], sd_src[NSL_D], sd_dst[NSL_D];
#pragma unroll
    for (int e = 0; e < NSL; e++) slide_entry(P, 0, MULTI ? SPLIT : P.ns, tid + e * NT, sl_src[e], sl_dst[e]);
#pragma unroll
    for (int e = 0; e < NSL_D; e++) { sd_src[e] = -1; sd_dst[e] = -1; if (MULTI) slide_entry(P, SPLIT, P.ns, tid + e * NT, sd_src[e], sd_dst[e]); }
    __syncthreads();
    if (P.tune) pstep = s_pstep;
    // register-resident taps of the plan's FIR stages (lane's tap split is ts = tid & 7)
    FirTaps ft;
    if constexpr (sizeof...(PLAN) > 0) {
        const double *cs = reinterpret_cast<const double *>(sm + P.coef_sm);
        int fi = 0;
        for (int s = 0; s < P.ns && fi < MAXFIR; s++) {
            if (P.st[s].type == 1) {
#pragma unroll
                for (int kk = 0; kk < FIR_KB; kk++) {
                    const double v = cs[P.st[s].coef + (tid & 7) * FIR_KB + kk];
                    if (fi == 0) ft.cf[0][kk] = v; else ft.cf[1][kk] = v;
                }
                fi++;
            }
        }
    }

    const cd *gin = P.in + (size_t)c * P.in_stride;
    cd *gout = P.out + (size_t)c * P.out_stride;
    const FStage &S0 = P.st[0];
    // stage 0's new samples: element k*NT + tid lands at pb0 + k*step0 (NT is a multiple of the pad unit)
    // (stage 0 is padded every 2*R0 elements when it is a half band, unpadded when it is a FIR)
    cd *pb0 = sm + S0.buf + phys(S0, S0.Ha + tid);
    constexpr int PU0 = PlanFirst<PLAN...>::code / 100 == 2 ? 4 * R0 : 2 * R0;      // stage 0's pad unit
    constexpr int STEP_PAD = NT + NT / PU0;
    const bool pad0 = S0.pu != 0;
    const bool tune = P.tune != 0;

    const int n_full = P.n_in / P.T0;
    const int rem = P.n_in - n_full * P.T0;
    constexpr bool FIXED = sizeof...(PLAN) > 0; // plan kernels always run T0 == 2048: no per-load guards
    const int nld = FIXED ? NLD : P.T0 / NT;    // loads per thread in a full chunk (uniform)

    // The next chunk's samples are fetched into registers while the current chunk is worked on:
    // up to NLD independent 16-byte loads per thread stay in flight across the whole cascade.
    cd nx[NLD];
    if (n_full > 0) {
#pragma unroll
        for (int k = 0; k < NLD; k++) if (FIXED || k < nld) nx[k] = gin[k * NT + tid];
    }
    int out_pos = 0;
    int deep_acc = 0;           // multi-rate plans: samples waiting in front of the first deep stage
    for (int ch = 0; ch < n_full; ch++) {
        long long *tr = (P.trace && tid == 0 && ch < 16) ? P.trace + ((size_t)c * 16 + ch) * 16 : nullptr;
        if (tr) tr[0] = clock64();
        // ---- commit the prefetched chunk behind stage 0's history, applying the NCO on the way
        if (tune) {
#pragma unroll
            for (int k = 0; k < NLD; k++) {
                if (FIXED || k < nld) {
                    const cd q = s_q[k];
                    const cd v = make_double2(fma(u.x, q.x, -u.y * q.y), fma(u.x, q.y, u.y * q.x));
                    nx[k] = make_double2(fma(nx[k].x, v.x, -nx[k].y * v.y), fma(nx[k].x, v.y, nx[k].y * v.x));
                }
            }
            u = make_double2(fma(u.x, pstep.x, -u.y * pstep.y), fma(u.x, pstep.y, u.y * pstep.x));
        }
        if (pad0) {
#pragma unroll
            for (int k = 0; k < NLD; k++) if (FIXED || k < nld) pb0[k * STEP_PAD] = nx[k];
        } else {
#pragma unroll
            for (int k = 0; k < NLD; k++) if (FIXED || k < nld) pb0[k * NT] = nx[k];
        }
        __syncthreads();
        if (tr) tr[1] = clock64();
        // ---- start fetching the next full chunk
        if (ch + 1 < n_full) {
            const cd *g1 = gin + (size_t)(ch + 1) * P.T0 + tid;
#pragma unroll
            for (int k = 0; k < NLD; k++) if (FIXED || k < nld) nx[k] = g1[k * NT];
        }
        if (tr) tr[2] = clock64();
        // ---- the cascade
        if constexpr (MULTI) {
            const int per = P.st[SPLIT - 1].n_out_full;
            cascade_c<NT, true, 0, SPLIT, 0, 0, PLAN...>(sm, P, P.T0, nullptr, tr, ft, deep_acc);
            deep_acc += per;
            if (deep_acc == per * P.deepk) {
                out_pos += cascade_c<NT, true, SPLIT, NS, 0, 0, PLAN...>(sm, P, deep_acc, gout + out_pos, tr, ft, 0);
                cd kd[NSL_D];
#pragma unroll
                for (int e = 0; e < NSL_D; e++) if (sd_src[e] >= 0) kd[e] = sm[sd_src[e]];
                __syncthreads();
#pragma unroll
                for (int e = 0; e < NSL_D; e++) if (sd_src[e] >= 0) sm[sd_dst[e]] = kd[e];
                deep_acc = 0;
            }
        } else if constexpr (NS > 0) {
            out_pos += cascade_c<NT, true, 0, NS, 0, 0, PLAN...>(sm, P, P.T0, gout + out_pos, tr, ft, 0);
        } else {
            for (int s = 0; s < P.ns; s++) {
                run_stage<NT>(sm, P, s, P.st[s].n_out_full, gout + out_pos);
                __syncthreads();
            }
            out_pos += P.st[P.ns - 1].n_out_full;
        }
        // ---- slide the histories: tail -> registers, barrier, registers -> front.  The writes are
        //      ordered against the next reads by the barrier that follows the next commit.
        cd keep[NSL];
#pragma unroll
        for (int e = 0; e < NSL; e++) if (sl_src[e] >= 0) keep[e] = sm[sl_src[e]];
        __syncthreads();
#pragma unroll
        for (int e = 0; e < NSL; e++) if (sl_src[e] >= 0) sm[sl_dst[e]] = keep[e];
        if (tr) tr[15] = clock64();
    }
    __syncthreads();
    // ---- ragged tail (at most one partial chunk), generic indexing
    int n_s = 0;
    if (rem > 0) {
        const cd *g1 = gin + (size_t)n_full * P.T0;
        for (int i = tid, k = 0; i < rem; i += NT, k++) {
            cd x = g1[i];
            if (P.tune) {
                const cd q = s_q[k];
                const cd v = make_double2(fma(u.x, q.x, -u.y * q.y), fma(u.x, q.y, u.y * q.x));
                x = make_double2(fma(x.x, v.x, -x.y * v.y), fma(x.x, v.y, x.y * v.x));
            }
            sm[S0.buf + phys(S0, S0.Ha + i)] = x;
        }
        __syncthreads();
        if constexpr (MULTI) {
            deep_acc += cascade_c<NT, false, 0, SPLIT, 0, 0, PLAN...>(sm, P, rem, nullptr, nullptr, ft, deep_acc);
        } else if constexpr (NS > 0) {
            cascade_c<NT, false, 0, NS, 0, 0, PLAN...>(sm, P, rem, gout + out_pos, nullptr, ft, 0);
        } else {
            int n_in = rem;
            for (int s = 0; s < P.ns; s++) {
                const int n_out = stage_out_count(P.st[s], n_in);
                run_stage<NT>(sm, P, s, n_out, gout + out_pos);
                n_in = n_out;
                __syncthreads();
            }
        }
        n_s = rem;
    }
    // multi-rate plans: whatever is still waiting in front of the deep stages goes through them now
    if constexpr (MULTI) {
        if (deep_acc > 0) cascade_c<NT, false, SPLIT, NS, 0, 0, PLAN...>(sm, P, deep_acc, gout + out_pos, nullptr, ft, 0);
    }
    // ---- hand the histories back: the last Hs inputs of every stage
    for (int s = 0; s < P.ns; s++) {
        const FStage &S = P.st[s];
        if (MULTI && s == SPLIT) n_s = deep_acc;
        cd *h = S.hout + (size_t)c * S.Hs;
        for (int i = tid; i < S.Hs; i += NT) h[i] = sm[S.buf + phys(S, n_s + (S.Ha - S.Hs) + i)];
        n_s = stage_out_count(S, n_s);
    }
}

// ---- Tail-warp variant of the plan kernels.  The stages behind the fourth half band see 128 samples or fewer per
// chunk: run by all 128 threads they are three barrier-separated phases of pure latency (window loads -> a 13-deep FMA
// chain -> three shuffle rounds), 2000 of a chunk's 6900 cycles for 12 % of its flops.  Here two extra warps (the "tail warps") own them:
// the four main warps run commit + stages 0 .. TS-1 of chunk n while the tail warp runs stages TS .. of chunk n-1 out
// of the other half of a double-buffered stage-TS input.  Hand-over by named barriers (PTX producer/consumer pattern):
// READY[p] = 2 + p (main arrives, tail waits), FREE[p] = 4 + p (tail arrives, main waits before refilling half p);
// barrier 1 is the main warps' own.  The stage-TS history is carried from the tail of one half to the front of the other.
// Same arithmetic in the same order as fused_decim_kernel: bit-identical outputs and state.
// 160 registers: two CTAs of six warps put three warps on every SM sub-partition (16384 registers each); at 168 they
// fill it to the last 256 registers and the one-warp-per-sub-partition CTAs of nco_advance_kernel (768 registers per
// warp) can only run by displacing a whole decimator CTA for the 0.36 ms they live.
#ifndef TW_MAXREG
#define TW_MAXREG 160
#endif
__device__ __forceinline__ void bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void bar_arrive(int id, int n) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); }

// ASYNC (round 2, last third): the next chunk does not wait in 64 registers per main-warp thread while the cascade runs.
// Each thread copies its sixteen samples straight from HBM into their (padded) places in stage 0's buffer with 16-byte
// cp.async (LDGSTS: per-thread destinations, so the pad every 2 R0 elements costs nothing), issued as soon as the first
// half band has consumed the buffer and its history has been slid; at the top of the next chunk the thread waits for
// its own group and applies the tuning phasor in place (it reads back only what it copied itself, so no barrier in
// between).  Costs one more 16-byte shared-memory read per sample; frees the registers for the half bands' own loads
// in flight.  Same arithmetic: bit-identical.  ASYNC = 2 is the same under a 128-register cap.
__device__ __forceinline__ void cp_async16(unsigned dst, const void *src)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

template <int ASYNC, int TS, int... PLAN>
__global__ void __maxnreg__(ASYNC == 2 ? 128 : TW_MAXREG) fused_decim_tw_kernel(const __grid_constant__ FusedParams P)
{
    constexpr int NT = 128, NTT = 64, NTA = NT + NTT, R0 = 8, NLD = 2 * R0, T0 = NLD * NT, NS = (int)sizeof...(PLAN);
    extern __shared__ double smem_raw[];
    cd *sm = reinterpret_cast<cd *>(smem_raw);
    const int c = blockIdx.x;
    const int tid = threadIdx.x;
    __shared__ cd s_pstep;
    __shared__ cd s_q[NLD];

    // the first chunk and the stage histories are requested before anything else: their latency (one HBM round trip
    // each; the histories were seven dependent round trips when loaded stage by stage) hides under the zero fill
    const int n_full = P.n_in / T0;
    const int rem = P.n_in - n_full * T0;
    const cd *gin = P.in + (size_t)c * P.in_stride;
    // Main warp lane `tid` carries sample `js` of every row of NT samples, rotated so that the eight lanes of a
    // quarter warp (one wavefront of a 16-byte store) start on a multiple of 8 elements of stage 0's padded buffer:
    // the buffer has one pad element every 2 R0 = 16, its origin is fixed by the window start of thread 0, and with
    // js = tid every other quarter warp straddled a pad and took two wavefronts (ncu: +50 % on the commit's stores,
    // 4 % of all shared-memory wavefronts of the kernel).
    const int js = (tid + ((8 - ((P.st[0].Ha + P.st[0].org) & 7)) & 7)) & (NT - 1);
    cd nx[NLD];
    if (!ASYNC && tid < NT && n_full > 0) {
#pragma unroll
        for (int k = 0; k < NLD; k++) nx[k] = gin[k * NT + js];
    }
    constexpr int NHL = 3;                      // history elements per thread (at most 3 x 192 = 576 >= the 512 the planner admits)
    cd hv[NHL]; int hdst[NHL];
#pragma unroll
    for (int e = 0; e < NHL; e++) {
        int idx = tid + e * NTA;
        hdst[e] = -1;
        for (int s = 0; s < P.ns; s++) {
            const FStage &S = P.st[s];
            if (idx >= 0 && idx < S.Hs) { hdst[e] = S.buf + phys(S, (S.Ha - S.Hs) + idx); hv[e] = S.hin[(size_t)c * S.Hs + idx]; }
            idx -= S.Hs;
        }
    }
    for (int i = tid; i < P.smem_cd; i += NTA) sm[i] = make_double2(0.0, 0.0);
    __syncthreads();
    constexpr int PU0 = PlanFirst<PLAN...>::code / 100 == 2 ? 4 * R0 : 2 * R0;      // stage 0's pad unit (component split: twice as wide)
    constexpr int STEP_PAD0 = NT + NT / PU0;
    unsigned pb0s = 0;
    if constexpr (ASYNC != 0) {
        const FStage &S0 = P.st[0];
        pb0s = (unsigned)__cvta_generic_to_shared(sm + S0.buf + phys(S0, S0.Ha + js));
        if (tid < NT && n_full > 0) {
#pragma unroll
            for (int k = 0; k < NLD; k++) cp_async16(pb0s + k * STEP_PAD0 * (int)sizeof(cd), gin + k * NT + js);
            cp_async_commit();
        }
    }
#pragma unroll
    for (int e = 0; e < NHL; e++) if (hdst[e] >= 0) sm[hdst[e]] = hv[e];
    {
        double *cs = reinterpret_cast<double *>(sm + P.coef_sm);
        for (int i = tid; i < P.ncoef; i += NTA) cs[i] = P.coef[i];
    }
    cd u = make_double2(1.0, 0.0), pstep = make_double2(1.0, 0.0);
    if (P.tune) {
        const double *nc = P.nco + (size_t)c * 8;
        if (tid == 0) s_pstep = nco_pow(nc, (unsigned long long)T0);
        if (tid < NLD) s_q[tid] = nco_pow(nc, (unsigned long long)tid * NT);
        if (tid < NT) u = cmul_rn(P.vstart[c], nco_pow(nc, P.n_base + (unsigned long long)js));
    }
    __syncthreads();
    const FStage &ST = P.st[TS];
    const int tws = P.tw_stride;
    cd *gout = P.out + (size_t)c * P.out_stride;
    FirTaps ft;
    auto load_taps = [&](int lane) {
        const double *cs = reinterpret_cast<const double *>(sm + P.coef_sm);
        int fi = 0;
        for (int s = 0; s < P.ns && fi < MAXFIR; s++) {
            if (P.st[s].type == 1) {
#pragma unroll
                for (int kk = 0; kk < FIR_KB; kk++) {
                    const double v = cs[P.st[s].coef + (lane & 7) * FIR_KB + kk];
                    if (fi == 0) ft.cf[0][kk] = v; else ft.cf[1][kk] = v;
                }
                fi++;
            }
        }
    };

    if (tid >= NT) {
        // ================= tail warp: stages TS .. NS-1, one chunk behind the main warps
        const int t = tid - NT;
        load_taps(t);
        constexpr int NSLT = 384 / NTT;         // in-place history slides of stages TS+1 .. (at most 384 elements)
        int ts_src[NSLT], ts_dst[NSLT];
#pragma unroll
        for (int e = 0; e < NSLT; e++) slide_entry(P, TS + 1, P.ns, t + e * NTT, ts_src[e], ts_dst[e]);
        int out_pos = 0;
        for (int ch = 0; ch < n_full; ch++) {
            const int p = ch & 1;
            bar_sync(2 + p, NTA);
            out_pos += cascade_x<NTT, 2, true, TS, NS, 0, 0, PLAN...>(sm, P, 0, gout + out_pos, ft, t, p * tws, 0);
            const cd *hs = sm + ST.buf + p * tws;
            cd *hd = sm + ST.buf + (p ^ 1) * tws;
            for (int i = t; i < ST.Ha; i += NTT) hd[phys(ST, i)] = hs[phys(ST, ST.n_full + i)];
            cd kd[NSLT];
#pragma unroll
            for (int e = 0; e < NSLT; e++) if (ts_src[e] >= 0) kd[e] = sm[ts_src[e]];
            group_sync<2>();
#pragma unroll
            for (int e = 0; e < NSLT; e++) if (ts_src[e] >= 0) sm[ts_dst[e]] = kd[e];
            group_sync<2>();
            if (ch + 2 < n_full) bar_arrive(4 + p, NTA);
        }
    } else {
        // ================= main warps: NCO + commit + stages 0 .. TS-1
        constexpr int NSL = 2;                  // four half bands: 4 x 48 history elements
        int sl_src[NSL], sl_dst[NSL];
#pragma unroll
        for (int e = 0; e < NSL; e++) slide_entry(P, ASYNC ? 1 : 0, TS, tid + e * NT, sl_src[e], sl_dst[e]);
        if (P.tune) pstep = s_pstep;
        const FStage &S0 = P.st[0];
        cd *pb0 = sm + S0.buf + phys(S0, S0.Ha + js);
        constexpr int STEP_PAD = NT + NT / PU0;
        const bool tune = P.tune != 0;
        // ASYNC: element i of stage 0's history is slid by the thread whose last cp.async of the next chunk lands on its
        // old place (logical Ha + T0 - Ha + i = row NLD-1, column NT - Ha + i), so the read precedes the overwrite in
        // program order (the store of the value read has issued before the copy does)
        int s0_src = -1, s0_dst = -1;
        if constexpr (ASYNC != 0) {
            const int i = js - (NT - S0.Ha);
            if (i >= 0 && i < S0.Ha) { s0_src = S0.buf + phys(S0, S0.n_full + i); s0_dst = S0.buf + phys(S0, i); }
        }
        for (int ch = 0; ch < n_full; ch++) {
            const int p = ch & 1;
            if constexpr (ASYNC != 0) {
                cp_async_wait_all();
                if (tune) {
#pragma unroll
                    for (int k = 0; k < NLD; k++) nx[k] = pb0[k * STEP_PAD];       // all loads first: the stores below must not fence them
#pragma unroll
                    for (int k = 0; k < NLD; k++) {
                        const cd q = s_q[k];
                        const cd v = make_double2(fma(u.x, q.x, -u.y * q.y), fma(u.x, q.y, u.y * q.x));
                        nx[k] = make_double2(fma(nx[k].x, v.x, -nx[k].y * v.y), fma(nx[k].x, v.y, nx[k].y * v.x));
                    }
#pragma unroll
                    for (int k = 0; k < NLD; k++) pb0[k * STEP_PAD] = nx[k];
                    u = make_double2(fma(u.x, pstep.x, -u.y * pstep.y), fma(u.x, pstep.y, u.y * pstep.x));
                }
                group_sync<1>();
                cascade_x<NT, 1, true, 0, 1, 0, 0, PLAN...>(sm, P, T0, nullptr, ft, tid, 0, 0);
                if (s0_src >= 0) sm[s0_dst] = sm[s0_src];
                if (ch + 1 < n_full) {
                    const cd *g1 = gin + (size_t)(ch + 1) * T0 + js;
#pragma unroll
                    for (int k = 0; k < NLD; k++) cp_async16(pb0s + k * STEP_PAD * (int)sizeof(cd), g1 + k * NT);
                    cp_async_commit();
                }
                cascade_x<NT, 1, true, 1, TS - 1, 0, 0, PLAN...>(sm, P, T0, nullptr, ft, tid, 0, 0);
            } else {
                if (tune) {
#pragma unroll
                    for (int k = 0; k < NLD; k++) {
                        const cd q = s_q[k];
                        const cd v = make_double2(fma(u.x, q.x, -u.y * q.y), fma(u.x, q.y, u.y * q.x));
                        nx[k] = make_double2(fma(nx[k].x, v.x, -nx[k].y * v.y), fma(nx[k].x, v.y, nx[k].y * v.x));
                    }
                    u = make_double2(fma(u.x, pstep.x, -u.y * pstep.y), fma(u.x, pstep.y, u.y * pstep.x));
                }
#pragma unroll
                for (int k = 0; k < NLD; k++) pb0[k * STEP_PAD] = nx[k];
                group_sync<1>();
                if (ch + 1 < n_full) {
                    const cd *g1 = gin + (size_t)(ch + 1) * T0 + js;
#pragma unroll
                    for (int k = 0; k < NLD; k++) nx[k] = g1[k * NT];
                }
                cascade_x<NT, 1, true, 0, TS - 1, 0, 0, PLAN...>(sm, P, T0, nullptr, ft, tid, 0, 0);
            }
            if (ch >= 2) bar_sync(4 + p, NTA);                  // the tail warp has finished with half p (chunk ch - 2)
            cascade_x<NT, 1, true, TS - 1, TS, 0, 0, PLAN...>(sm, P, 0, nullptr, ft, tid, 0, p * tws);
            bar_arrive(2 + p, NTA);
            cd keep[NSL];
#pragma unroll
            for (int e = 0; e < NSL; e++) if (sl_src[e] >= 0) keep[e] = sm[sl_src[e]];
            group_sync<1>();
#pragma unroll
            for (int e = 0; e < NSL; e++) if (sl_src[e] >= 0) sm[sl_dst[e]] = keep[e];
        }
    }
    __syncthreads();
    if (tid >= NT) return;
    // the stage-TS history sits at the front of half (n_full & 1): bring it to half 0, where the code below expects it
    if (n_full & 1) {
        for (int i = tid; i < ST.Ha; i += NT) sm[ST.buf + phys(ST, i)] = sm[ST.buf + tws + phys(ST, i)];
    }
    group_sync<1>();
    // ---- ragged tail (at most one partial chunk): every stage on the main warps, generic indexing
    int n_s = 0;
    if (rem > 0) {
        load_taps(tid);
        const cd *g1 = P.in + (size_t)c * P.in_stride + (size_t)n_full * T0;
        const FStage &S0 = P.st[0];
        for (int i = js, k = 0; i < rem; i += NT, k++) {
            cd x = g1[i];
            if (P.tune) {
                const cd q = s_q[k];
                const cd v = make_double2(fma(u.x, q.x, -u.y * q.y), fma(u.x, q.y, u.y * q.x));
                x = make_double2(fma(x.x, v.x, -x.y * v.y), fma(x.x, v.y, x.y * v.x));
            }
            sm[S0.buf + phys(S0, S0.Ha + i)] = x;
        }
        group_sync<1>();
        cascade_x<NT, 1, false, 0, NS, 0, 0, PLAN...>(sm, P, rem, gout + n_full * P.st[NS - 1].n_out_full, ft, tid, 0, 0);
        n_s = rem;
    }
    for (int s = 0; s < P.ns; s++) {
        const FStage &S = P.st[s];
        cd *h = S.hout + (size_t)c * S.Hs;
        for (int i = tid; i < S.Hs; i += NT) h[i] = sm[S.buf + phys(S, n_s + (S.Ha - S.Hs) + i)];
        n_s = stage_out_count(S, n_s);
    }
}


// ---- Three-group pipeline (round 2, last third).  The tail-warp kernel's four main warps still walk through commit ->
// half band 0 -> half band 1 -> half band 2 as barrier-separated phases, the last two at R = 4 / R = 2 outputs per thread
// (7.25 / 12.5 shared-memory loads per output).  Here the chunk pipeline has three stations, each one chunk behind the
// one in front, all busy at the same time:
//   group A, 4 warps: tuning phasor + commit + half band 0 (R = 8)            of chunk n      -> half n & 1 of stage 1's input
//   group B, 2 warps: half band 1 (R = 8) and half band 2 (R = 4)             of chunk n - 1  -> half of stage 3's input
//   group C, 2 warps: half band 3, FIR, half band, FIR (the former tail warps) of chunk n - 2  -> HBM
// so every half band in front of the tail streams at 4.6 - 7.25 loads per output and the FP64 work per warp is even
// (12.6 / 14 / 10.5 % of a chunk).  Hand-over by named barriers: READY / FREE per half and per pair of groups
// (A-B: 8+p / 10+p, 192 threads; B-C: 2+p / 4+p, 128 threads), each group keeps a barrier of its own (1, 7, 6).
// Eight warps per CTA, two CTAs per SM under a 128-register cap.  Same arithmetic in the same order: bit-identical.
template <int... PLAN>
__global__ void __launch_bounds__(256, 2) fused_decim_p3_kernel(const __grid_constant__ FusedParams P)
{
    constexpr int NT = 128, NTB = 64, NTT = 64, NTA = NT + NTB + NTT, R0 = 8, NLD = 2 * R0, T0 = NLD * NT, NS = (int)sizeof...(PLAN), TS = 3;
    extern __shared__ double smem_raw[];
    cd *sm = reinterpret_cast<cd *>(smem_raw);
    const int c = blockIdx.x;
    const int tid = threadIdx.x;
    __shared__ cd s_pstep;
    __shared__ cd s_q[NLD];

    const int n_full = P.n_in / T0;
    const int rem = P.n_in - n_full * T0;
    const cd *gin = P.in + (size_t)c * P.in_stride;
    const int js = (tid + ((8 - ((P.st[0].Ha + P.st[0].org) & 7)) & 7)) & (NT - 1);      // see fused_decim_tw_kernel
    cd nx[NLD];
    if (tid < NT && n_full > 0) {
#pragma unroll
        for (int k = 0; k < NLD; k++) nx[k] = gin[k * NT + js];
    }
    constexpr int NHL = 2;                      // history elements per thread (2 x 256 = the 512 the planner admits)
    cd hv[NHL]; int hdst[NHL];
#pragma unroll
    for (int e = 0; e < NHL; e++) {
        int idx = tid + e * NTA;
        hdst[e] = -1;
        for (int s = 0; s < P.ns; s++) {
            const FStage &S = P.st[s];
            if (idx >= 0 && idx < S.Hs) { hdst[e] = S.buf + phys(S, (S.Ha - S.Hs) + idx); hv[e] = S.hin[(size_t)c * S.Hs + idx]; }
            idx -= S.Hs;
        }
    }
    for (int i = tid; i < P.smem_cd; i += NTA) sm[i] = make_double2(0.0, 0.0);
    __syncthreads();
#pragma unroll
    for (int e = 0; e < NHL; e++) if (hdst[e] >= 0) sm[hdst[e]] = hv[e];
    {
        double *cs = reinterpret_cast<double *>(sm + P.coef_sm);
        for (int i = tid; i < P.ncoef; i += NTA) cs[i] = P.coef[i];
    }
    cd u = make_double2(1.0, 0.0), pstep = make_double2(1.0, 0.0);
    if (P.tune) {
        const double *nc = P.nco + (size_t)c * 8;
        if (tid == NT) s_pstep = nco_pow(nc, (unsigned long long)T0);
        if (tid >= NT + NTB && tid < NT + NTB + NLD) s_q[tid - NT - NTB] = nco_pow(nc, (unsigned long long)(tid - NT - NTB) * NT);
        if (tid < NT) u = cmul_rn(P.vstart[c], nco_pow(nc, P.n_base + (unsigned long long)js));
    }
    __syncthreads();
    const FStage &S1 = P.st[1];
    const FStage &ST = P.st[TS];
    const int abs_ = P.ab_stride, tws = P.tw_stride;
    cd *gout = P.out + (size_t)c * P.out_stride;
    FirTaps ft;
    auto load_taps = [&](int lane) {
        const double *cs = reinterpret_cast<const double *>(sm + P.coef_sm);
        int fi = 0;
        for (int s = 0; s < P.ns && fi < MAXFIR; s++) {
            if (P.st[s].type == 1) {
#pragma unroll
                for (int kk = 0; kk < FIR_KB; kk++) {
                    const double v = cs[P.st[s].coef + (lane & 7) * FIR_KB + kk];
                    if (fi == 0) ft.cf[0][kk] = v; else ft.cf[1][kk] = v;
                }
                fi++;
            }
        }
    };

    if (tid >= NT + NTB) {
        // ================= group C: stages TS .. NS-1, two chunks behind group A
        const int t = tid - NT - NTB;
        load_taps(t);
        constexpr int NSLT = 384 / NTT;
        int ts_src[NSLT], ts_dst[NSLT];
#pragma unroll
        for (int e = 0; e < NSLT; e++) slide_entry(P, TS + 1, P.ns, t + e * NTT, ts_src[e], ts_dst[e]);
        int out_pos = 0;
        for (int ch = 0; ch < n_full; ch++) {
            const int p = ch & 1;
            long long *tr = (P.trace && t == 0 && ch < 16) ? P.trace + ((size_t)c * 16 + ch) * 16 : nullptr;
            if (tr) tr[8] = clock64();
            bar_sync(2 + p, NTB + NTT);
            if (tr) tr[9] = clock64();
            out_pos += cascade_x<NTT, 2, true, TS, NS, 0, 0, PLAN...>(sm, P, 0, gout + out_pos, ft, t, p * tws, 0);
            if (tr) tr[10] = clock64();
            const cd *hs = sm + ST.buf + p * tws;
            cd *hd = sm + ST.buf + (p ^ 1) * tws;
            for (int i = t; i < ST.Ha; i += NTT) hd[phys(ST, i)] = hs[phys(ST, ST.n_full + i)];
            cd kd[NSLT];
#pragma unroll
            for (int e = 0; e < NSLT; e++) if (ts_src[e] >= 0) kd[e] = sm[ts_src[e]];
            group_sync<2>();
#pragma unroll
            for (int e = 0; e < NSLT; e++) if (ts_src[e] >= 0) sm[ts_dst[e]] = kd[e];
            group_sync<2>();
            if (ch + 2 < n_full) bar_arrive(4 + p, NTB + NTT);
            if (tr) tr[11] = clock64();
        }
    } else if (tid >= NT) {
        // ================= group B: half bands 1 and 2, one chunk behind group A
        const int t = tid - NT;
        int b_src, b_dst;
        slide_entry(P, 2, 3, t, b_src, b_dst);               // stage 2's history (48 elements) slides in place
        for (int ch = 0; ch < n_full; ch++) {
            const int p = ch & 1;
            long long *tr = (P.trace && t == 0 && ch < 16) ? P.trace + ((size_t)c * 16 + ch) * 16 : nullptr;
            if (tr) tr[4] = clock64();
            bar_sync(8 + p, NT + NTB);
            if (tr) tr[5] = clock64();
            cascade_x<NTB, 3, true, 1, 2, 0, 0, PLAN...>(sm, P, 0, nullptr, ft, t, p * abs_, 0);
            {   // stage 1's history: from the end of this half to the front of the other
                const cd *hs = sm + S1.buf + p * abs_;
                cd *hd = sm + S1.buf + (p ^ 1) * abs_;
                for (int i = t; i < S1.Ha; i += NTB) hd[phys(S1, i)] = hs[phys(S1, S1.n_full + i)];
            }
            if (ch + 2 < n_full) bar_arrive(10 + p, NT + NTB);
            if (tr) tr[6] = clock64();
            if (ch >= 2) bar_sync(4 + p, NTB + NTT);           // group C has finished with half p of stage 3's input
            if (tr) tr[7] = clock64();
            cascade_x<NTB, 3, true, 2, 3, 0, 0, PLAN...>(sm, P, 0, nullptr, ft, t, 0, p * tws);
            bar_arrive(2 + p, NTB + NTT);
            cd keep = make_double2(0.0, 0.0);
            if (b_src >= 0) keep = sm[b_src];
            group_sync<3>();
            if (b_src >= 0) sm[b_dst] = keep;
            if (tr) tr[12] = clock64();
        }
    } else {
        // ================= group A: NCO + commit + half band 0
        int a_src, a_dst;
        slide_entry(P, 0, 1, tid, a_src, a_dst);
        if (P.tune) pstep = s_pstep;
        const FStage &S0 = P.st[0];
        cd *pb0 = sm + S0.buf + phys(S0, S0.Ha + js);
        constexpr int STEP_PAD = NT + NT / (2 * R0);
        const bool tune = P.tune != 0;
        for (int ch = 0; ch < n_full; ch++) {
            const int p = ch & 1;
            long long *tr = (P.trace && tid == 0 && ch < 16) ? P.trace + ((size_t)c * 16 + ch) * 16 : nullptr;
            if (tr) tr[0] = clock64();
            if (tune) {
#pragma unroll
                for (int k = 0; k < NLD; k++) {
                    const cd q = s_q[k];
                    const cd v = make_double2(fma(u.x, q.x, -u.y * q.y), fma(u.x, q.y, u.y * q.x));
                    nx[k] = make_double2(fma(nx[k].x, v.x, -nx[k].y * v.y), fma(nx[k].x, v.y, nx[k].y * v.x));
                }
                u = make_double2(fma(u.x, pstep.x, -u.y * pstep.y), fma(u.x, pstep.y, u.y * pstep.x));
            }
#pragma unroll
            for (int k = 0; k < NLD; k++) pb0[k * STEP_PAD] = nx[k];
            group_sync<1>();
            if (ch + 1 < n_full) {
                const cd *g1 = gin + (size_t)(ch + 1) * T0 + js;
#pragma unroll
                for (int k = 0; k < NLD; k++) nx[k] = g1[k * NT];
            }
            if (tr) tr[1] = clock64();
            if (ch >= 2) bar_sync(10 + p, NT + NTB);            // group B has finished with half p of stage 1's input
            if (tr) tr[2] = clock64();
            cascade_x<NT, 1, true, 0, 1, 0, 0, PLAN...>(sm, P, T0, nullptr, ft, tid, 0, p * abs_);
            bar_arrive(8 + p, NT + NTB);
            if (tr) tr[3] = clock64();
            cd keep = make_double2(0.0, 0.0);
            if (a_src >= 0) keep = sm[a_src];
            group_sync<1>();
            if (a_src >= 0) sm[a_dst] = keep;
        }
    }
    __syncthreads();
    if (tid >= NT) return;
    // the histories of the double-buffered stages sit at the front of half (n_full & 1): bring them to half 0
    if (n_full & 1) {
        for (int i = tid; i < S1.Ha; i += NT) sm[S1.buf + phys(S1, i)] = sm[S1.buf + abs_ + phys(S1, i)];
        for (int i = tid; i < ST.Ha; i += NT) sm[ST.buf + phys(ST, i)] = sm[ST.buf + tws + phys(ST, i)];
    }
    group_sync<1>();
    // ---- ragged tail (at most one partial chunk): every stage on group A, generic indexing
    int n_s = 0;
    if (rem > 0) {
        load_taps(tid);
        const cd *g1 = P.in + (size_t)c * P.in_stride + (size_t)n_full * T0;
        const FStage &S0 = P.st[0];
        for (int i = js, k = 0; i < rem; i += NT, k++) {
            cd x = g1[i];
            if (P.tune) {
                const cd q = s_q[k];
                const cd v = make_double2(fma(u.x, q.x, -u.y * q.y), fma(u.x, q.y, u.y * q.x));
                x = make_double2(fma(x.x, v.x, -x.y * v.y), fma(x.x, v.y, x.y * v.x));
            }
            sm[S0.buf + phys(S0, S0.Ha + i)] = x;
        }
        group_sync<1>();
        cascade_x<NT, 1, false, 0, NS, 0, 0, PLAN...>(sm, P, rem, gout + n_full * P.st[NS - 1].n_out_full, ft, tid, 0, 0);
        n_s = rem;
    }
    for (int s = 0; s < P.ns; s++) {
        const FStage &S = P.st[s];
        cd *h = S.hout + (size_t)c * S.Hs;
        for (int i = tid; i < S.Hs; i += NT) h[i] = sm[S.buf + phys(S, n_s + (S.Ha - S.Hs) + i)];
        n_s = stage_out_count(S, n_s);
    }
}

struct FusedDecimator {
    FusedParams P;
    size_t smem_bytes = 0;
    bool attr_set = false;
};

static bool stage_fusable(const BatchFilter *f)
{
    if (f->kind == QC_C_DECIM2_HB45) return true;
    if (f->kind == QC_C_DECIMATE) return f->decim == 1 || f->decim == 2 || f->decim == 3 || f->decim == 5;
    return false;
}

size_t RxChain::fusable_prefix(size_t limit)
{
    if (!fd) fd = new FusedDecimator();
    // leading run of half-band / FIR-decimate stages (quisk_process_decimate, then the demodulator's own)
    size_t n = 0;
    int ncoef = 0, hist = 0;
    long dtot = 1;
    while (n < limit && n < cst.size() && n < (size_t)MAXST && stage_fusable(cst[n])) {
        const BatchFilter *f = cst[n];
        long d = f->kind == QC_C_DECIM2_HB45 ? 2 : f->decim;
        int kp = 0, ha = 48;
        if (f->kind == QC_C_DECIMATE) { kp = ((f->nTaps + 8 * FIR_KB - 1) / (8 * FIR_KB)) * (8 * FIR_KB); ha = kp + 8; }
        // the chunk must be a multiple of both the total decimation and the CTA width
        long l = dtot * d;
        long lcm = l;
        while (lcm % FNT) lcm += l;
        if (lcm > 2048 || ncoef + kp > MAXCOEF || hist + ha > 512) break;
        dtot = l; ncoef += kp; hist += ha; n++;
    }
    return n;
}

int RxChain::run_fused_decimator(size_t n_stages, const cd *in, long in_stride, int count, cd *out, long out_stride, int *n_out, cudaStream_t strm)
{
    FusedParams &P = fd->P;
    const int ns = (int)n_stages;
    // total decimation and chunk size
    int dtot = 1;
    for (int s = 0; s < ns; s++) dtot *= (cst[s]->kind == QC_C_DECIM2_HB45 ? 2 : cst[s]->decim);
    int lcm = dtot;
    while (lcm % FNT) lcm += dtot;
    int T0 = (fused_chunk / lcm) * lcm;
    if (T0 <= 0) T0 = lcm;
    if (T0 > 2048) T0 = (2048 / lcm) * lcm;
    const int NT = fused_threads == 256 ? 256 : 128;
    if (T0 % NT) T0 = ((T0 / NT) * NT > 0 && ((T0 / NT) * NT) % lcm == 0) ? (T0 / NT) * NT : T0;
    if (T0 % NT) { set_error("fused decimator: chunk %d is not a multiple of the CTA width %d", T0, NT); return QC_EINVAL; }
    P.ns = ns; P.T0 = T0;
    P.in = in; P.in_stride = in_stride; P.n_in = count; P.out = out; P.out_stride = out_stride;
    P.nco = d_nco; P.vstart = d_v[vcur]; P.n_base = 0; P.tune = tune ? 1 : 0;
    P.trace = d_trace;
    // multi-rate split: the first stage that sees <= 128 samples per chunk, and everything after it, runs
    // once every `deepk` chunks (plan kernels only)
    int split = ns, deepk = 1;
    if (fused_plans && fused_deepk > 1 && NT == 128 && T0 == 2048) {
        int ci = T0;
        for (int s = 0; s < ns; s++) {
            if (ci <= 128 && s > 0) { split = s; deepk = fused_deepk; break; }
            ci /= (cst[s]->kind == QC_C_DECIM2_HB45 ? 2 : cst[s]->decim);
        }
    }
    int codes[MAXST];
    size_t sh = 0;
    // component-split half bands (hb_stage_split): plan kernels at the full chunk only
    // tail-warp kernel: stages from index tw_ts on belong to the two tail warps (option value 2, 3 or 4; 1 = default)
    const int tw_ts = fused_tailwarp >= 2 && fused_tailwarp <= 4 ? fused_tailwarp : 3;
    bool use_tw = fused_tailwarp && fused_plans && NT == 128 && T0 == 2048 && split == ns && ns > 4 && (!d_trace || fused_p3);
    // (split = 2, the default, also takes the four-stage 192 kS/s plan: measured 0.530 -> 0.509 ms per launch at 1024 receivers)
    bool use_split = !use_tw && (fused_split == 1 || (fused_split >= 2 && ns == 4)) && fused_plans && NT == 128 && T0 == 2048 && split == ns;
    // tail-warp kernel with component-split half bands behind the first one (stage 0 keeps complex lanes: its commit
    // layout is built around the pad every 16 elements)
    bool tw_split = use_tw && fused_split && (tw_ts == 3 || (tw_ts == 4 && fused_split == 3));
    // three-group pipeline (fused_decim_p3_kernel): half bands 1 and 2 on a group of their own
    bool use_p3 = use_tw && fused_p3 && tw_ts == 3;
    if (use_p3) tw_split = false;
    // stage descriptors for a given multi-rate split (no filter state is touched here)
    auto build = [&](int split, int deepk) -> int {
    P.deepk = deepk;
    int off = 0, coff = 0, chunk_in = T0;
    for (int s = 0; s < ns; s++) {
        BatchFilter *f = cst[s];
        FStage &S = P.st[s];
        memset(&S, 0, sizeof(S));
        if (s == split) chunk_in *= deepk;          // deep stages see deepk chunks' worth per run
        S.Hs = f->H;
        S.hin = (const cd *)f->d_hist[f->cur];
        S.hout = (cd *)f->d_hist[f->cur ^ 1];
        if (f->kind == QC_C_DECIM2_HB45) {
            S.type = 0; S.D = 2; S.nTaps = 43; S.Ha = 48;
            S.u0 = 1 - f->phase;
            const int nout = chunk_in / 2;
            const int nts = ((use_tw && s >= tw_ts) || (use_p3 && s >= 1)) ? 64 : NT;          // the tail warps (and group B) are 64 threads
            S.R = nout > 4 * nts ? 8 : (nout > 2 * nts ? 4 : 2);
            if (fused_min_r > S.R) S.R = fused_min_r;
            if (use_split || (tw_split && (s > 0 || fused_split == 3))) { S.split = 1; S.R *= 2; }
            S.pu = 2 * S.R;
            S.magic = (unsigned)((0x100000000ULL + S.pu - 1) / S.pu);
            // window start of thread 0 (logical Ha + u0 - 42) must land on a pad boundary
            const int ws = S.Ha + S.u0 - 42;
            S.org = (S.pu - (ws % S.pu)) % S.pu;
            const int qmax = S.Ha + chunk_in + S.org + 2 * S.R + (use_p3 ? 16 : 64);      // p3: two more buffer halves have to fit
            S.buf_len = qmax + qmax / S.pu + 8;
            const int q0 = ws + S.org;
            S.p0 = q0 + q0 / S.pu;
        } else {
            S.type = 1; S.D = f->decim; S.nTaps = f->nTaps; S.TS = 8;
            S.Kpad = ((f->nTaps + S.TS * FIR_KB - 1) / (S.TS * FIR_KB)) * (S.TS * FIR_KB);
            S.Ha = S.Kpad + 8;
            S.u0 = f->decim - 1 - f->phase;
            S.R = FIR_R; S.pu = 0; S.magic = 0; S.org = 0;
            {   // plan kernels: outputs per thread so that one round of NT threads covers the chunk
                const int no = chunk_in / S.D;
                const int nts = (use_tw && s >= tw_ts) ? 64 : NT;
                int r = (no * 8 + nts - 1) / nts;
                S.Rplan = r >= 4 ? 4 : (r >= 2 ? 2 : 1);
            }
            S.coef = coff;
            for (int k = 0; k < S.Kpad; k++) P.coef[coff + k] = k < f->nTaps ? f->h_coef[k] : 0.0;
            coff += S.Kpad;
            // the last thread group may read R*D samples past the chunk when n_out is not a multiple of R
            S.buf_len = S.Ha + chunk_in + S.D * FIR_R + 16;
        }
        S.buf = off;
        off += S.buf_len;
        if (use_tw && s == tw_ts) { P.tw_stride = S.buf_len; off += S.buf_len; }
        if (use_p3 && s == 1) { P.ab_stride = S.buf_len; off += S.buf_len; }
        S.n_full = chunk_in;
        S.n_out_full = chunk_in / S.D;
        chunk_in = chunk_in / S.D;
        codes[s] = S.type ? 100 + S.Rplan * 10 + S.D : (S.split ? 200 + (S.R / 2) * 10 + 2 : S.R * 10 + 2);
    }
    P.scratch = off;
    {
        int tot = 0, deep = 0;
        for (int s = 0; s < ns; s++) { if (s < split) tot += P.st[s].Ha; else deep += P.st[s].Ha; }
        if (tot > 512 || deep > 384) { set_error("fused decimator: %d/%d history elements exceed the slide capacity", tot, deep); return QC_EINVAL; }
    }
    P.coef_sm = off; P.ncoef = coff;
    off += (coff + 1) / 2 + 1;
    P.smem_cd = off;
    sh = (size_t)off * sizeof(cd);
    if (sh > 226 * 1024) { set_error("fused decimator: %zu bytes of shared memory needed", sh); return QC_EINVAL; }
    return QC_OK;
    };
    // the multi-rate plans that have a compiled instantiation
    auto multi_plan_exists = [&]() {
        static const int plans[4][MAXST] = {{82, 42, 22, 22, 142}, {82, 42, 22, 22, 142, 22, 142},
                                            {82, 42, 22, 22, 142, 22, 22, 122}, {82, 42, 22, 22, 142, 142}};
        static const int lens[4] = {5, 7, 8, 6};
        for (int p = 0; p < 4; p++) {
            if (lens[p] != ns) continue;
            bool ok = true;
            for (int i = 0; i < ns; i++) ok = ok && codes[i] == plans[p][i];
            if (ok) return true;
        }
        return false;
    };
    int rcb;
    if (use_split) {
        static const int plans[5][MAXST] = {{282, 242, 222, 222, 142}, {282, 242, 222, 222, 142, 222, 112},
                                            {282, 242, 222, 222, 142, 222, 222, 112}, {282, 242, 222, 222, 142, 122},
                                            {282, 142, 222, 142}};
        static const int lens[5] = {5, 7, 8, 6, 4};
        rcb = build(ns, 1);
        bool found = false;
        for (int p = 0; p < 5 && rcb == QC_OK; p++) {
            if (lens[p] != ns) continue;
            bool ok = true;
            for (int i = 0; i < ns; i++) ok = ok && codes[i] == plans[p][i];
            found = found || ok;
        }
        if (!found) use_split = false;
    }
    if (use_tw) {
        // the half bands in front, sized for 128 main or 64 tail threads
        const int head[3][4] = {{82, 42, 42, 22}, {82, 42, 22, 22}, {82, 42, 22, 22}};        // tw_ts = 2, 3, 4
        static const int tails[4][MAXST] = {{142}, {142, 22, 122}, {142, 22, 22, 112}, {142, 142}};
        static const int lens[4] = {5, 7, 8, 6};
        rcb = build(ns, 1);
        bool found = false;
        for (int p = 0; p < 4 && rcb == QC_OK; p++) {
            if (lens[p] != ns) continue;
            bool ok = true;
            for (int i = 0; i < ns; i++) {
                int want = i < 4 ? head[tw_ts - 2][i] : tails[p][i - 4];
                if (tw_split && (i > 0 || fused_split == 3) && want < 100) want += 200;
                if (use_p3 && i < 4) want = i < 2 ? 82 : (i == 2 ? 42 : 22);
                ok = ok && codes[i] == want;
            }
            found = found || ok;
        }
        for (int s = 0; s < ns; s++) if (P.st[s].type == 1 && P.st[s].Kpad != 8 * FIR_KB) found = false;
        if (!found) { use_tw = false; tw_split = false; use_p3 = false; P.tw_stride = 0; P.ab_stride = 0; }
    }
    if (use_tw || use_split) {
    } else if (split != ns) {
        rcb = build(split, deepk);
        if (rcb != QC_OK || split != 4 || !multi_plan_exists()) { split = ns; deepk = 1; }
    }
    if (split == ns && !use_split && !use_tw) { rcb = build(ns, 1); if (rcb != QC_OK) return rcb; }
    // host-side phase bookkeeping, same formulas as BatchFilter::run
    int n = count;
    for (int s = 0; s < ns; s++) {
        BatchFilter *f = cst[s];
        const int no = f->count_out(n, 0);
        if (f->kind == QC_C_DECIM2_HB45) f->phase = (f->phase + n) & 1;
        else f->phase = (f->phase + n) % f->decim;
        f->cur ^= 1;
        n = no;
    }
    *n_out = n;
    const int R0 = P.st[0].type == 0 ? (P.st[0].split ? P.st[0].R / 2 : P.st[0].R) : 2;       // complex outputs per thread (pair)
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (timing) {
        QC_CUDA(cudaEventCreate(&e0)); QC_CUDA(cudaEventCreate(&e1));
        QC_CUDA(cudaEventRecord(e0, strm));
    }
    // plan code per stage: type*100 + R*10 + D
    bool plan_ok = fused_plans && T0 == 2 * R0 * NT;
    for (int s = 0; s < ns; s++) if (P.st[s].type == 1 && P.st[s].Kpad != 8 * FIR_KB) plan_ok = false;
    auto is_plan = [&](int sp, std::initializer_list<int> pl, int nt = 128) {
        if (!plan_ok || (int)pl.size() != ns || sp != split || nt != NT) return false;
        int i = 0;
        for (int c : pl) if (codes[i++] != c) return false;
        return true;
    };
#define QC_LAUNCH(...) do { \
        static bool optin[64] = {}; \
        int dev_ = 0; cudaGetDevice(&dev_); dev_ &= 63; \
        if (!optin[dev_]) { QC_CUDA(cudaFuncSetAttribute(fused_decim_kernel<__VA_ARGS__>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024)); optin[dev_] = true; } \
        fused_name = "fused_decim_kernel<" #__VA_ARGS__ ">"; \
        fused_decim_kernel<__VA_ARGS__><<<C, NT, sh, strm>>>(P); } while (0)
#define QC_LAUNCH_TW(...) do { \
        static bool optin[64] = {}; \
        int dev_ = 0; cudaGetDevice(&dev_); dev_ &= 63; \
        if (!optin[dev_]) { QC_CUDA(cudaFuncSetAttribute(fused_decim_tw_kernel<__VA_ARGS__>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024)); optin[dev_] = true; } \
        fused_name = "fused_decim_tw_kernel<" #__VA_ARGS__ ">"; \
        fused_decim_tw_kernel<__VA_ARGS__><<<C, 192, sh, strm>>>(P); } while (0)
    // tail-warp kernels (default): the stages behind the fourth half band run on a fifth warp, one chunk behind
    if (use_tw) {
#define QC_TW_PLANS0(AS, TS, H0, H1, H2, H3) \
        if (ns == 5) QC_LAUNCH_TW(AS, TS, H0, H1, H2, H3, 142); \
        else if (ns == 7) QC_LAUNCH_TW(AS, TS, H0, H1, H2, H3, 142, H3, 122); \
        else if (ns == 8) QC_LAUNCH_TW(AS, TS, H0, H1, H2, H3, 142, H3, H3, 112); \
        else QC_LAUNCH_TW(AS, TS, H0, H1, H2, H3, 142, 142)
#define QC_TW_PLANS(AS, TS, H1, H2, H3) QC_TW_PLANS0(AS, TS, 82, H1, H2, H3)
        if (use_p3) {
#define QC_LAUNCH_P3(...) do { \
        static bool optin[64] = {}; \
        int dev_ = 0; cudaGetDevice(&dev_); dev_ &= 63; \
        if (!optin[dev_]) { QC_CUDA(cudaFuncSetAttribute(fused_decim_p3_kernel<__VA_ARGS__>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024)); optin[dev_] = true; } \
        fused_name = "fused_decim_p3_kernel<" #__VA_ARGS__ ">"; \
        fused_decim_p3_kernel<__VA_ARGS__><<<C, 256, sh, strm>>>(P); } while (0)
            if (ns == 5) QC_LAUNCH_P3(82, 82, 42, 22, 142);
            else if (ns == 7) QC_LAUNCH_P3(82, 82, 42, 22, 142, 22, 122);
            else if (ns == 8) QC_LAUNCH_P3(82, 82, 42, 22, 142, 22, 22, 112);
            else QC_LAUNCH_P3(82, 82, 42, 22, 142, 142);
#undef QC_LAUNCH_P3
        }
        else if (tw_split && fused_split == 3 && tw_ts == 4) { QC_TW_PLANS0(0, 4, 282, 242, 222, 222); }
        else if (tw_ts == 2) { QC_TW_PLANS(0, 2, 42, 42, 22); } else if (tw_ts == 4) { QC_TW_PLANS(0, 4, 42, 22, 22); }
        else if (tw_split && fused_split == 3) { QC_TW_PLANS0(0, 3, 282, 242, 222, 222); }
        else if (tw_split) { QC_TW_PLANS(0, 3, 242, 222, 222); }
        else if (fused_async == 1) { QC_TW_PLANS(1, 3, 42, 22, 22); } else if (fused_async == 2) { QC_TW_PLANS(2, 3, 42, 22, 22); }
        else { QC_TW_PLANS(0, 3, 42, 22, 22); }
#undef QC_TW_PLANS
#undef QC_TW_PLANS0
    }
    // single-rate plans (every stage every chunk)
    else if (is_plan(5, {82, 42, 22, 22, 142})) QC_LAUNCH(128, 8, 2, 5, 82, 42, 22, 22, 142);                        // 1.536 MS/s -> 48 k
    else if (is_plan(7, {82, 42, 22, 22, 142, 22, 112})) QC_LAUNCH(128, 8, 2, 7, 82, 42, 22, 22, 142, 22, 112);   // ... -> 12 k (SSB)
    else if (is_plan(8, {82, 42, 22, 22, 142, 22, 22, 112})) QC_LAUNCH(128, 8, 2, 8, 82, 42, 22, 22, 142, 22, 22, 112);   // ... -> 6 k (CW)
    else if (is_plan(6, {82, 42, 22, 22, 142, 122})) QC_LAUNCH(128, 8, 2, 6, 82, 42, 22, 22, 142, 122);           // ... -> 24 k (AM)
    // the same four with component-split half bands (default)
    else if (is_plan(5, {282, 242, 222, 222, 142})) QC_LAUNCH(128, 8, 2, 5, 282, 242, 222, 222, 142);
    else if (is_plan(7, {282, 242, 222, 222, 142, 222, 112})) QC_LAUNCH(128, 8, 2, 7, 282, 242, 222, 222, 142, 222, 112);
    else if (is_plan(8, {282, 242, 222, 222, 142, 222, 222, 112})) QC_LAUNCH(128, 8, 2, 8, 282, 242, 222, 222, 142, 222, 222, 112);
    else if (is_plan(6, {282, 242, 222, 222, 142, 122})) QC_LAUNCH(128, 8, 2, 6, 282, 242, 222, 222, 142, 122);
    // 192 kS/s -> 12 k (SSB at the north star's target rate): HB45 + FIR98/2 to 48 k, HB45 + FIR98/2 to 12 k
    else if (is_plan(4, {282, 142, 222, 142})) QC_LAUNCH(128, 8, 2, 4, 282, 142, 222, 142);
    else if (is_plan(4, {82, 142, 22, 142})) QC_LAUNCH(128, 8, 2, 4, 82, 142, 22, 142);
    // half-size chunks (1024 samples): 8 loads in flight per thread instead of 16, half the shared memory, three CTAs per SM
    else if (is_plan(7, {42, 22, 22, 22, 122, 22, 112}, 128)) QC_LAUNCH(128, 4, 3, 7, 42, 22, 22, 22, 122, 22, 112);
    else if (is_plan(5, {42, 22, 22, 22, 122}, 128)) QC_LAUNCH(128, 4, 3, 5, 42, 22, 22, 22, 122);
    // 256-thread CTAs, two per SM under a 128-register cap (16 warps per SM)
    else if (is_plan(7, {42, 22, 22, 22, 122, 22, 112}, 256)) QC_LAUNCH(256, 4, 2, 7, 42, 22, 22, 22, 122, 22, 112);
    else if (is_plan(5, {42, 22, 22, 22, 122}, 256)) QC_LAUNCH(256, 4, 2, 5, 42, 22, 22, 22, 122);
    // multi-rate plans: stages from index 4 on run once per 4 chunks
    else if (is_plan(4, {82, 42, 22, 22, 142})) QC_LAUNCH(128, 8, 2, 4, 82, 42, 22, 22, 142);
    else if (is_plan(4, {82, 42, 22, 22, 142, 22, 142})) QC_LAUNCH(128, 8, 2, 4, 82, 42, 22, 22, 142, 22, 142);
    else if (is_plan(4, {82, 42, 22, 22, 142, 22, 22, 122})) QC_LAUNCH(128, 8, 2, 4, 82, 42, 22, 22, 142, 22, 22, 122);
    else if (is_plan(4, {82, 42, 22, 22, 142, 142})) QC_LAUNCH(128, 8, 2, 4, 82, 42, 22, 22, 142, 142);
    else if (NT == 256) {
        if (R0 == 4) { if (fused_dense) QC_LAUNCH(256, 4, 2, 0); else QC_LAUNCH(256, 4, 1, 0); }
        else if (R0 == 2) { if (fused_dense) QC_LAUNCH(256, 2, 2, 0); else QC_LAUNCH(256, 2, 1, 0); }
        else { set_error("fused decimator: unsupported stage-0 blocking %d at 256 threads", R0); return QC_EINVAL; }
    } else {
        if (R0 == 8) { if (fused_dense) QC_LAUNCH(128, 8, 4, 0); else QC_LAUNCH(128, 8, 2, 0); }
        else if (R0 == 4) { if (fused_dense) QC_LAUNCH(128, 4, 4, 0); else QC_LAUNCH(128, 4, 2, 0); }
        else { if (fused_dense) QC_LAUNCH(128, 2, 4, 0); else QC_LAUNCH(128, 2, 2, 0); }
    }
#undef QC_LAUNCH
#undef QC_LAUNCH_TW
    count_launch();
    QC_CUDA_LAUNCH();
    if (timing) { QC_CUDA(cudaEventRecord(e1, strm)); timed.push_back(std::make_pair(e0, e1)); }
    return QC_OK;
}

int RxChain::reset_fused() { return QC_OK; }

void RxChain::release_fused()
{
    if (fd) { delete fd; fd = nullptr; }
}

}  // namespace qc
