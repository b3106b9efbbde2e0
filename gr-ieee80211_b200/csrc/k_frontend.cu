// k_frontend.cu -- everything before the per-symbol loop:
//   k_presiso : the presiso hier block (examples/presiso.grc:35-229): x[n-16]*conj(x[n]) -> moving
//               sum 48, |x|^2 -> moving sum 64, preac = |c| / p.  Data-parallel, HBM streaming:
//               8 B read + 4 B (preac) [+ 8 B preconj] written per sample; a warp sweeps its segment with the
//               sliding tree in registers.  Also writes the 1-bit-per-sample threshold bitmap.
//   k_trigger : trigger FSM over one array (staged entry point, lib/trigger_impl.cc:59-117)
//   k_detect  : one thread per item: trigger FSM -> sync -> signal (phy_serial.cuh::detect_item)      } frontend_mode 1; the
//   k_header  : one thread per frame: format detection, SIG fields, channel estimate                  } default kernels are
//   k_header2 : the same for the 2-antenna block (signal2 / demod2)                                   } in k_frontend_w.cu
#include "common.cuh"
#include "phy_serial.cuh"

namespace {

using c8b::cf;

#ifndef C8B_PRESISO_CTAS
#define C8B_PRESISO_CTAS 6
#endif
constexpr int PWARPS = 4;           // warps per CTA, each sweeping its own segment
constexpr int PSEG = 160;           // rows (32 samples) of output per warp segment
constexpr int PDEPTH = 4;           // rows of iq in flight per warp = rows per unrolled group
constexpr unsigned PFULL = 0xffffffffu;
static_assert(PSEG % PDEPTH == 0, "a segment is a whole number of row groups");

struct f3 { float x, y, z; };       // (re, im) of the lag-16 product sum and |x|^2 sum
__device__ __forceinline__ f3 add3(f3 a, f3 b) { return { __fadd_rn(a.x, b.x), __fadd_rn(a.y, b.y), __fadd_rn(a.z, b.z) }; }

// per-lane all-ones / all-zeros words: "this lane sends its previous-row value" for the lags 1, 2, 4, 8, 16.  Kept opaque
// so the selects stay single LOP3s on registers instead of a compare + FSEL per row.
struct PMasks { uint32_t m1, m2, m4, m8, m16; };
__device__ __forceinline__ uint32_t opaque(uint32_t m) { asm volatile("" : "+r"(m)); return m; }
__device__ __forceinline__ float selm(uint32_t m, float a, float b)          // m ? a : b
{
    return __uint_as_float((__float_as_uint(a) & m) | (__float_as_uint(b) & ~m));
}
// value `lag` samples back along the stream: lane - lag of this row, or of the previous row for the first `lag` lanes
__device__ __forceinline__ f3 back3(f3 cur, f3 prev, uint32_t m, int src)
{
    return { __shfl_sync(PFULL, selm(m, prev.x, cur.x), src), __shfl_sync(PFULL, selm(m, prev.y, cur.y), src),
             __shfl_sync(PFULL, selm(m, prev.z, cur.z), src) };
}
// value 16 samples back: the partner lane (lane ^ 16) of this row for the upper half-warp, of the previous row for the lower
__device__ __forceinline__ float back16(float cur, float prev, uint32_t m16) { return __shfl_xor_sync(PFULL, selm(m16, prev, cur), 16); }

struct PState {                     // previous row: samples, products, tree levels; |x|^2 s16 two rows back
    float2 xp; f3 vp, s2p, s4p, s8p, s16p; float pw2p;
};

// one row: lane l <-> sample i = 32 r + l.  pa / pc / pm point at this row's outputs (pm: the row's bitmap word).
template <bool OUT, bool CONJ, bool MASK>
__device__ __forceinline__ void presiso_row(PState& S, const float2 c, const int i, const int n, const PMasks& M, const int lane,
                                            float* __restrict__ pa, float2* __restrict__ pc, uint32_t* __restrict__ pm)
{
    const float2 dl = make_float2(back16(c.x, S.xp.x, M.m16), back16(c.y, S.xp.y, M.m16));      // delay(16)
    f3 v;
    v.x = __fadd_rn(__fmul_rn(dl.x, c.x), __fmul_rn(dl.y, c.y));                                // in0 * conj(in1)
    v.y = __fsub_rn(__fmul_rn(dl.y, c.x), __fmul_rn(dl.x, c.y));
    v.z = __fadd_rn(__fmul_rn(c.x, c.x), __fmul_rn(c.y, c.y));                                  // complex_to_mag_squared
    if (i < 16) { v.x = 0.f; v.y = 0.f; }
    const f3 s2 = add3(back3(v, S.vp, M.m1, (lane - 1) & 31), v);
    const f3 s4 = add3(back3(s2, S.s2p, M.m2, (lane - 2) & 31), s2);
    const f3 s8 = add3(back3(s4, S.s4p, M.m4, (lane - 4) & 31), s4);
    const f3 s16 = add3(back3(s8, S.s8p, M.m8, (lane - 8) & 31), s8);
    if (OUT) {
        const f3 a1 = { back16(s16.x, S.s16p.x, M.m16), back16(s16.y, S.s16p.y, M.m16), back16(s16.z, S.s16p.z, M.m16) };
        const float pw3 = back16(S.s16p.z, S.pw2p, M.m16);
        const float cr = __fadd_rn(__fadd_rn(S.s16p.x, a1.x), s16.x), ci = __fadd_rn(__fadd_rn(S.s16p.y, a1.y), s16.y);
        const float pw = __fadd_rn(__fadd_rn(pw3, S.s16p.z), __fadd_rn(a1.z, s16.z));
        const float mag = __fsqrt_rn(__fadd_rn(__fmul_rn(cr, cr), __fmul_rn(ci, ci)));          // complex_to_mag
        const float ac = __fdiv_rn(mag, pw);                                                    // divide_ff
        const bool in = i < n;
        if (in) {
            *pa = ac;
            if (CONJ) *pc = make_float2(cr, ci);
        }
        if (MASK) {                                             // threshold bitmap for the trigger scan (lib/trigger_impl.cc:79)
            const uint32_t m = __ballot_sync(PFULL, in && ac > 0.3f);
            if (lane == 0 && in) *pm = m;
        }
    }
    S.xp = c; S.vp = v; S.s2p = s2; S.s4p = s4; S.s8p = s8; S.pw2p = S.s16p.z; S.s16p = s16;
}

// The moving sums are evaluated as a fixed sliding tree so every output is a pure function of the 64
// (80) samples before it, independent of tiling: s2[n]=v[n-1]+v[n]; s4[n]=s2[n-2]+s2[n]; s8; s16;
// sum48[n]=(s16[n-32]+s16[n-16])+s16[n]; sum64[n]=(s16[n-48]+s16[n-32])+(s16[n-16]+s16[n]).
// A warp sweeps a segment of one item row by row (lane l <-> sample 32 r + l): each sample is loaded once, coalesced,
// and every lag is a register of the previous row or one shuffle -- no shared memory, no barrier.  A segment that
// does not start the item first runs one group of history rows (3 are needed: x[i-16] of the products, the 15-sample
// tree, the 48 lag) without output.
template <bool CONJ, bool MASK>
__global__ void __launch_bounds__(PWARPS * 32, C8B_PRESISO_CTAS)
k_presiso(const float2* __restrict__ iq, const int64_t* __restrict__ off, const int32_t* __restrict__ len, int nitems, int nseg,
          int64_t outBase, float* __restrict__ preac, float2* __restrict__ preconj, uint32_t* __restrict__ mask, int maskStride)
{
    const int lane = threadIdx.x & 31;
    const int64_t wid = (int64_t)blockIdx.x * PWARPS + (threadIdx.x >> 5);
    const int item = (int)(wid / nseg), seg = (int)(wid % nseg);
    if (item >= nitems) return;
    const int n = len[item];
    const int r0 = seg * PSEG;                                  // first output row
    if (r0 * 32 >= n) return;
    const int nEnd = min(n, (r0 + PSEG) * 32);                  // one past the last sample this warp needs
    const int rStart = r0 > 0 ? r0 - PDEPTH : 0;
    int i = rStart * 32 + lane;
    const float2* __restrict__ x = iq + off[item];
    const int iLast = nEnd - 1;                                 // loads past the end re-read this sample: those lanes store nothing
    const int64_t ob = off[item] - outBase;
    float* __restrict__ pa = preac + ob + i;
    float2* __restrict__ pc = CONJ ? preconj + ob + i : nullptr;
    uint32_t* __restrict__ pm = MASK ? mask + (size_t)item * maskStride + rStart : nullptr;
    PMasks M;
    M.m1 = opaque(lane >= 31 ? ~0u : 0u); M.m2 = opaque(lane >= 30 ? ~0u : 0u); M.m4 = opaque(lane >= 28 ? ~0u : 0u);
    M.m8 = opaque(lane >= 24 ? ~0u : 0u); M.m16 = opaque(lane >= 16 ? ~0u : 0u);

    const float2 z2 = make_float2(0.f, 0.f);
    float2 q[PDEPTH];
#pragma unroll
    for (int d = 0; d < PDEPTH; d++) q[d] = __ldg(x + min(i + 32 * d, iLast));
    const f3 z3 = { 0.f, 0.f, 0.f };
    PState S = { z2, z3, z3, z3, z3, z3, 0.f };
    if (r0 > 0) {                                               // history rows of a later segment
        float2 nq[PDEPTH];                                      // the next group's rows, requested back to back (1 KB contiguous)
#pragma unroll
        for (int d = 0; d < PDEPTH; d++) nq[d] = __ldg(x + min(i + 32 * (d + PDEPTH), iLast));
#pragma unroll
        for (int d = 0; d < PDEPTH; d++) presiso_row<false, CONJ, MASK>(S, q[d], i + 32 * d, n, M, lane, nullptr, nullptr, nullptr);
#pragma unroll
        for (int d = 0; d < PDEPTH; d++) q[d] = nq[d];
        i += 32 * PDEPTH; pa += 32 * PDEPTH;
        if (CONJ) pc += 32 * PDEPTH;
        if (MASK) pm += PDEPTH;
    }
    for (; i - lane < nEnd; i += 32 * PDEPTH) {                 // causal sums: samples at or past n never reach a stored output
        float2 nq[PDEPTH];
#pragma unroll
        for (int d = 0; d < PDEPTH; d++) nq[d] = __ldg(x + min(i + 32 * (d + PDEPTH), iLast));
#pragma unroll
        for (int d = 0; d < PDEPTH; d++)
            presiso_row<true, CONJ, MASK>(S, q[d], i + 32 * d, n, M, lane, pa + 32 * d, CONJ ? pc + 32 * d : nullptr, MASK ? pm + d : nullptr);
#pragma unroll
        for (int d = 0; d < PDEPTH; d++) q[d] = nq[d];
        pa += 32 * PDEPTH;
        if (CONJ) pc += 32 * PDEPTH;
        if (MASK) pm += PDEPTH;
    }
}

__global__ void k_trigger(const float* __restrict__ preac, int64_t n, uint8_t* __restrict__ out)
{
    if (blockIdx.x || threadIdx.x) return;
    c8b::TrigState ts;
    c8b::trig_reset(ts);
    for (int64_t i = 0; i < n; i++) out[i] = c8b::trig_step(ts, preac[i]);
}

__global__ void __launch_bounds__(64)
k_detect(const c8b_lut* __restrict__ lut, const float2* __restrict__ iq, const int64_t* __restrict__ off,
         const int32_t* __restrict__ len, int nitems, int itemBase, int maxf, int64_t outBase, const float* __restrict__ preac,
         const uint32_t* __restrict__ mask, int maskStride, c8b_frame* __restrict__ frames, float2* __restrict__ chan,
         c8b_scan* __restrict__ scans)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nitems) return;
    // records and channels are written in place (device global memory): frames[i*maxf ..], chan[i*maxf*64 ..]
    c8b::detect_item(lut, reinterpret_cast<const cf*>(iq + off[i]), preac + (off[i] - outBase), len[i], itemBase + i, maxf,
                     frames + (size_t)i * maxf, reinterpret_cast<cf*>(chan + (size_t)i * maxf * 64),
                     mask ? mask + (size_t)i * maskStride : nullptr, scans ? scans + i : nullptr);
}

struct RotSrc {
    const cf* x;      // sample 0 of the signal block's output = item sample sync_idx + 224
    float rad;
    int nsamp;
    __device__ cf operator()(int k) const
    {
        if (k >= nsamp) return c8b::mk(0.f, 0.f);                   // S_PAD: 320 samples the reference never writes
        return c8b::cmul(x[k], c8b::cis(c8b::fmul((float)(k + 224), rad)));   // lib/signal_impl.cc:164-192
    }
};

__global__ void __launch_bounds__(64)
k_header(const c8b_lut* __restrict__ lut, const float2* __restrict__ iq, const int64_t* __restrict__ off, int nslots, int maxf,
         int mupos, c8b_frame* __restrict__ frames, const float2* __restrict__ chan, float2* __restrict__ hinv,
         int64_t llrStride, float* __restrict__ llr)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;       // frame slot; its item is i / maxf
    if (i >= nslots) return;
    c8b_frame f = frames[i];
    f.llr_off = (int64_t)i * llrStride;
    if (f.status == C8B_ST_OK) {
        cf hl[64], hv[64];
        for (int k = 0; k < 64; k++) { const float2 c = chan[(size_t)i * 64 + k]; hl[k] = c8b::mk(c.x, c.y); }
        RotSrc rot;
        rot.x = reinterpret_cast<const cf*>(iq + off[i / maxf]) + f.sync_idx + 224;
        rot.rad = f.rad; rot.nsamp = f.nsamp;
        f.status = c8b::demod_header(lut, rot, f.nsamp, f.l_mcs, f.l_len, hl, mupos, &f, hv);
        for (int k = 0; k < 64; k++) hinv[(size_t)i * 64 + k] = make_float2(hv[k].re, hv[k].im);
        if (f.status == C8B_ST_OK && (int64_t)f.total > llrStride) f.status = C8B_ST_OVERFLOW;
        if (f.status == C8B_ST_NDP && llr && llrStride >= 256) {
            // tag "mu2x1chan" (lib/demod_impl.cc:238-249): the 2 x 64 time samples of the two VHT-LTFs that the sounding
            // branch of nonLegacyChanEstimate keeps (:391-394); a one-stream NDP never fills them (zeros here)
            float* o = llr + f.llr_off;
            for (int k = 0; k < 128; k++) {
                const cf c = f.nss != 1 ? rot(240 + C8B_SYM_SHIFT + (k & 63) + (k >> 6) * 80) : c8b::mk(0.f, 0.f);
                o[2 * k] = c.re; o[2 * k + 1] = c.im;
            }
        }
    }
    frames[i] = f;
}

// 2-antenna header states (demod2): per frame hinv (1-stream frames), w2 (2-stream frames: 256 ZF weights + 8 pilot refs)
__global__ void __launch_bounds__(64)
k_header2(const c8b_lut* __restrict__ lut, const float2* __restrict__ iq0, const float2* __restrict__ iq1,
          const int64_t* __restrict__ off, int nslots, int maxf, c8b_frame* __restrict__ frames, const float2* __restrict__ chan,
          float2* __restrict__ hinv, float2* __restrict__ w2, int64_t llrStride, int mmse)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nslots) return;
    c8b_frame f = frames[i];
    f.llr_off = (int64_t)i * llrStride;
    if (f.status == C8B_ST_OK) {
        cf hl[64], hv[64];
        for (int k = 0; k < 64; k++) { const float2 c = chan[(size_t)i * 64 + k]; hl[k] = c8b::mk(c.x, c.y); hv[k] = c8b::mk(0.f, 0.f); }
        RotSrc r0, r1;
        r0.x = reinterpret_cast<const cf*>(iq0 + off[i / maxf]) + f.sync_idx + 224; r0.rad = f.rad; r0.nsamp = f.nsamp;
        r1 = r0; r1.x = reinterpret_cast<const cf*>(iq1 + off[i / maxf]) + f.sync_idx + 224;
        f.status = c8b::demod_header2(lut, r0, r1, f.nsamp, f.l_mcs, f.l_len, hl, &f, hv, reinterpret_cast<cf*>(w2 + (size_t)i * 264),
                                       c8b::mmse_sigma2(mmse, f.snr, f.rssi));
        for (int k = 0; k < 64; k++) hinv[(size_t)i * 64 + k] = make_float2(hv[k].re, hv[k].im);
        if (f.status == C8B_ST_OK && (int64_t)f.total > llrStride) f.status = C8B_ST_OVERFLOW;
    }
    frames[i] = f;
}

}  // namespace

void c8b_launch_header2(const c8b_lut* lut, const float2* iq0, const float2* iq1, const int64_t* d_off, int nitems, int maxf,
                        int mmse, c8b_frame* frames, const float2* chan, float2* hinv, float2* w2, int64_t llrStride, cudaStream_t st)
{
    if (nitems <= 0) return;
    k_header2<<<(nitems * maxf + 63) / 64, 64, 0, st>>>(lut, iq0, iq1, d_off, nitems * maxf, maxf, frames, chan, hinv, w2, llrStride, mmse);
}

// sc16 -> fc32: the widening UHD's converter applies before the reference sees gr_complex (int16 * (1 / 32768), exact in
// float32).  Four samples per thread: 16 bytes in, 32 bytes out.
namespace {
__global__ void __launch_bounds__(256)
k_sc16_to_fc32(const short2* __restrict__ in, float2* __restrict__ out, int64_t n)
{
    const float sc = 1.0f / 32768.0f;
    const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (i + 4 <= n && ((reinterpret_cast<uintptr_t>(in + i) & 15) == 0)) {
        const int4 v = *reinterpret_cast<const int4*>(in + i);
        const short2 a = *reinterpret_cast<const short2*>(&v.x), b = *reinterpret_cast<const short2*>(&v.y);
        const short2 c = *reinterpret_cast<const short2*>(&v.z), d = *reinterpret_cast<const short2*>(&v.w);
        float4* o = reinterpret_cast<float4*>(out + i);
        o[0] = make_float4(a.x * sc, a.y * sc, b.x * sc, b.y * sc);
        o[1] = make_float4(c.x * sc, c.y * sc, d.x * sc, d.y * sc);
    } else {
        for (int64_t k = i; k < n && k < i + 4; k++) { const short2 a = in[k]; out[k] = make_float2(a.x * sc, a.y * sc); }
    }
}
}  // namespace

// Completion flag in mapped host memory: the last thing an op enqueues.  The host thread then polls that word instead of
// calling cudaStreamSynchronize -- with one thread per block (GNU Radio's scheduler) the waits of five blocks inside the
// driver queue up behind each other, a load from pinned memory does not.
namespace {
__global__ void k_flag(volatile uint32_t* flag, uint32_t seq)
{
    __threadfence_system();
    *flag = seq;
}
}  // namespace
void c8b_launch_flag(uint32_t* d_flag, uint32_t seq, cudaStream_t st) { k_flag<<<1, 1, 0, st>>>(d_flag, seq); }

int c8b_wait_flag(const volatile uint32_t* h_flag, uint32_t seq, cudaStream_t st)
{
    for (unsigned spins = 0;; spins++) {
        if (*h_flag == seq) return 0;
        if ((spins & 0xffff) == 0xffff) {                             // every ~65k polls: has the stream died / drained without the flag?
            const cudaError_t e = cudaStreamQuery(st);
            if (e == cudaSuccess) return *h_flag == seq ? 0 : -1;
            if (e != cudaErrorNotReady) return -1;
        }
#if defined(__x86_64__)
        __builtin_ia32_pause();
#endif
    }
}

void c8b_launch_sc16_to_fc32(const short2* in, float2* out, int64_t n, cudaStream_t st)
{
    if (n <= 0) return;
    const int64_t threads = (n + 3) / 4;
    k_sc16_to_fc32<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(in, out, n);
}

void c8b_launch_presiso(const float2* iq, const int64_t* d_off, const int32_t* d_len, int nitems, int maxLen, int64_t outBase,
                        float* preac, float2* preconj, uint32_t* mask, int maskStride, cudaStream_t st)
{
    if (nitems <= 0 || maxLen <= 0) return;
    const int nseg = (maxLen + PSEG * 32 - 1) / (PSEG * 32);
    const int64_t warps = (int64_t)nitems * nseg;
    const unsigned grid = (unsigned)((warps + PWARPS - 1) / PWARPS);
#define C8B_PRESISO(CJ, MK) k_presiso<CJ, MK><<<grid, PWARPS * 32, 0, st>>>(iq, d_off, d_len, nitems, nseg, outBase, preac, preconj, mask, maskStride)
    if (preconj) { if (mask) C8B_PRESISO(true, true); else C8B_PRESISO(true, false); }
    else         { if (mask) C8B_PRESISO(false, true); else C8B_PRESISO(false, false); }
#undef C8B_PRESISO
}

void c8b_launch_trigger(const float* preac, int64_t n, uint8_t* out, cudaStream_t st) { k_trigger<<<1, 32, 0, st>>>(preac, n, out); }

void c8b_launch_detect(const c8b_lut* lut, const float2* iq, const int64_t* d_off, const int32_t* d_len, int nitems, int itemBase,
                       int maxf, int64_t outBase, const float* preac, const uint32_t* mask, int maskStride, c8b_frame* frames,
                       float2* chan, c8b_scan* scans, cudaStream_t st)
{
    if (nitems <= 0) return;
    k_detect<<<(nitems + 63) / 64, 64, 0, st>>>(lut, iq, d_off, d_len, nitems, itemBase, maxf, outBase, preac, mask, maskStride, frames, chan,
                                                  scans);
}

void c8b_launch_header(const c8b_lut* lut, const float2* iq, const int64_t* d_off, int nitems, int maxf, int mupos, c8b_frame* frames,
                       const float2* chan, float2* hinv, int64_t llrStride, float* llr, cudaStream_t st)
{
    if (nitems <= 0) return;
    k_header<<<(nitems * maxf + 63) / 64, 64, 0, st>>>(lut, iq, d_off, nitems * maxf, maxf, mupos, frames, chan, hinv, llrStride, llr);
}
