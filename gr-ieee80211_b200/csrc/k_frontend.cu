// k_frontend.cu -- everything before the per-symbol loop:
//   k_presiso : the presiso hier block (examples/presiso.grc:35-229): x[n-16]*conj(x[n]) -> moving
//               sum 48, |x|^2 -> moving sum 64, preac = |c| / p.  Data-parallel, HBM streaming:
//               8 B read + 4 B (preac) [+ 8 B preconj] written per sample.
//   k_trigger : trigger FSM over one array (staged entry point, lib/trigger_impl.cc:59-117)
//   k_detect  : one thread per item: trigger FSM -> sync -> signal (phy_serial.cuh::detect_item)
//   k_header  : one thread per frame: format detection, SIG fields, channel estimate
#include "common.cuh"
#include "phy_serial.cuh"

namespace {

using c8b::cf;

constexpr int PWARPS = 4;           // warps per CTA, each sweeping its own segment
constexpr int PSEG = 160;           // rows (32 samples) of output per warp segment
constexpr int PDEPTH = 4;           // rows of iq in flight per warp
constexpr unsigned PFULL = 0xffffffffu;

struct f3 { float x, y, z; };       // (re, im) of the lag-16 product sum and |x|^2 sum
__device__ __forceinline__ f3 add3(f3 a, f3 b) { return { __fadd_rn(a.x, b.x), __fadd_rn(a.y, b.y), __fadd_rn(a.z, b.z) }; }

// value `lag` samples back along the stream: lane - lag of this row, or of the previous row for the first `lag` lanes
__device__ __forceinline__ float back(float cur, float prev, int lane, int lag)
{
    return __shfl_sync(PFULL, lane >= 32 - lag ? prev : cur, (lane - lag) & 31);
}
__device__ __forceinline__ f3 back3(f3 cur, f3 prev, int lane, int lag)
{
    return { back(cur.x, prev.x, lane, lag), back(cur.y, prev.y, lane, lag), back(cur.z, prev.z, lane, lag) };
}
// value 16 samples back: the partner lane (lane ^ 16) of this row for the upper half-warp, of the previous row for the lower
__device__ __forceinline__ float back16(float cur, float prev, int lane)
{
    return __shfl_xor_sync(PFULL, lane < 16 ? cur : prev, 16);
}

// The moving sums are evaluated as a fixed sliding tree so every output is a pure function of the 64
// (80) samples before it, independent of tiling: s2[n]=v[n-1]+v[n]; s4[n]=s2[n-2]+s2[n]; s8; s16;
// sum48[n]=(s16[n-32]+s16[n-16])+s16[n]; sum64[n]=(s16[n-48]+s16[n-32])+(s16[n-16]+s16[n]).
// A warp sweeps a segment of one item row by row (lane l <-> sample 32 r + l): each sample is loaded once, coalesced,
// and every lag is a register of the previous row or one shuffle -- no shared memory, no barrier.  A segment that
// does not start the item first runs 3 rows of history (x[i-16] of the products + the 15-sample tree + the 48 lag).
__global__ void __launch_bounds__(PWARPS * 32)
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
    const int rEnd = min(r0 + PSEG, (n + 31) >> 5);
    const int rStart = max(0, r0 - 3);
    const float2* __restrict__ x = iq + off[item];
    const int64_t ob = off[item] - outBase;
    uint32_t* __restrict__ mrow = mask ? mask + (size_t)item * maskStride : nullptr;

    auto fetch = [&](int r) {
        const int i = r * 32 + lane;
        return (r < rEnd && i < n) ? __ldg(x + i) : make_float2(0.f, 0.f);
    };
    float2 q[PDEPTH];
#pragma unroll
    for (int d = 0; d < PDEPTH; d++) q[d] = fetch(rStart + d);

    const f3 z3 = { 0.f, 0.f, 0.f };
    float2 xp = make_float2(0.f, 0.f);                          // previous row: samples, products, tree levels
    f3 vp = z3, s2p = z3, s4p = z3, s8p = z3, s16p = z3;
    float pw2p = 0.f;                                           // |x|^2 s16 two rows back
    for (int rb = rStart; rb < rEnd; rb += PDEPTH) {
#pragma unroll
        for (int d = 0; d < PDEPTH; d++) {
            const int r = rb + d;                               // rows past rEnd (at most PDEPTH - 1) are zeros, not stored
            const float2 c = q[d];
            q[d] = fetch(r + PDEPTH);
            const int i = r * 32 + lane;
            const float2 dl = make_float2(back16(c.x, xp.x, lane), back16(c.y, xp.y, lane));     // delay(16)
            f3 v;
            v.x = __fadd_rn(__fmul_rn(dl.x, c.x), __fmul_rn(dl.y, c.y));                        // in0 * conj(in1)
            v.y = __fsub_rn(__fmul_rn(dl.y, c.x), __fmul_rn(dl.x, c.y));
            v.z = __fadd_rn(__fmul_rn(c.x, c.x), __fmul_rn(c.y, c.y));                          // complex_to_mag_squared
            if (i < 16) { v.x = 0.f; v.y = 0.f; }
            const f3 s2 = add3(back3(v, vp, lane, 1), v);
            const f3 s4 = add3(back3(s2, s2p, lane, 2), s2);
            const f3 s8 = add3(back3(s4, s4p, lane, 4), s4);
            const f3 s16 = add3(back3(s8, s8p, lane, 8), s8);
            const f3 a1 = { back16(s16.x, s16p.x, lane), back16(s16.y, s16p.y, lane), back16(s16.z, s16p.z, lane) };
            const float pw3 = back16(s16p.z, pw2p, lane);
            const float cr = __fadd_rn(__fadd_rn(s16p.x, a1.x), s16.x), ci = __fadd_rn(__fadd_rn(s16p.y, a1.y), s16.y);
            const float pw = __fadd_rn(__fadd_rn(pw3, s16p.z), __fadd_rn(a1.z, s16.z));
            xp = c; vp = v; s2p = s2; s4p = s4; s8p = s8; pw2p = s16p.z; s16p = s16;
            if (r >= r0 && r < rEnd) {                                                          // warp-uniform
                const float mag = __fsqrt_rn(__fadd_rn(__fmul_rn(cr, cr), __fmul_rn(ci, ci)));  // complex_to_mag
                const float ac = __fdiv_rn(mag, pw);                                            // divide_ff
                const bool in = i < n;
                if (in) {
                    preac[ob + i] = ac;
                    if (preconj) preconj[ob + i] = make_float2(cr, ci);
                }
                if (mrow) {                                     // threshold bitmap for the trigger scan (lib/trigger_impl.cc:79)
                    const uint32_t m = __ballot_sync(PFULL, in && ac > 0.3f);
                    if (lane == 0) mrow[r] = m;
                }
            }
        }
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
         const uint32_t* __restrict__ mask, int maskStride, c8b_frame* __restrict__ frames, float2* __restrict__ chan)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nitems) return;
    // records and channels are written in place (device global memory): frames[i*maxf ..], chan[i*maxf*64 ..]
    c8b::detect_item(lut, reinterpret_cast<const cf*>(iq + off[i]), preac + (off[i] - outBase), len[i], itemBase + i, maxf,
                     frames + (size_t)i * maxf, reinterpret_cast<cf*>(chan + (size_t)i * maxf * 64),
                     mask ? mask + (size_t)i * maskStride : nullptr);
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
         int64_t llrStride)
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
    }
    frames[i] = f;
}

// 2-antenna header states (demod2): per frame hinv (1-stream frames), w2 (2-stream frames: 256 ZF weights + 8 pilot refs)
__global__ void __launch_bounds__(64)
k_header2(const c8b_lut* __restrict__ lut, const float2* __restrict__ iq0, const float2* __restrict__ iq1,
          const int64_t* __restrict__ off, int nslots, int maxf, c8b_frame* __restrict__ frames, const float2* __restrict__ chan,
          float2* __restrict__ hinv, float2* __restrict__ w2, int64_t llrStride)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nslots) return;
    c8b_frame f = frames[i];
    f.llr_off = (int64_t)i * llrStride;
    if (f.status == C8B_ST_OK) {
        cf hl[64], hv[64];
        for (int k = 0; k < 64; k++) { const float2 c = chan[(size_t)i * 64 + k]; hl[k] = c8b::mk(c.x, c.y); }
        RotSrc r0, r1;
        r0.x = reinterpret_cast<const cf*>(iq0 + off[i / maxf]) + f.sync_idx + 224; r0.rad = f.rad; r0.nsamp = f.nsamp;
        r1 = r0; r1.x = reinterpret_cast<const cf*>(iq1 + off[i / maxf]) + f.sync_idx + 224;
        f.status = c8b::demod_header2(lut, r0, r1, f.nsamp, f.l_mcs, f.l_len, hl, &f, hv, reinterpret_cast<cf*>(w2 + (size_t)i * 264));
        for (int k = 0; k < 64; k++) hinv[(size_t)i * 64 + k] = make_float2(hv[k].re, hv[k].im);
        if (f.status == C8B_ST_OK && (int64_t)f.total > llrStride) f.status = C8B_ST_OVERFLOW;
    }
    frames[i] = f;
}

}  // namespace

void c8b_launch_header2(const c8b_lut* lut, const float2* iq0, const float2* iq1, const int64_t* d_off, int nitems, int maxf,
                        c8b_frame* frames, const float2* chan, float2* hinv, float2* w2, int64_t llrStride, cudaStream_t st)
{
    if (nitems <= 0) return;
    k_header2<<<(nitems * maxf + 63) / 64, 64, 0, st>>>(lut, iq0, iq1, d_off, nitems * maxf, maxf, frames, chan, hinv, w2, llrStride);
}

void c8b_launch_presiso(const float2* iq, const int64_t* d_off, const int32_t* d_len, int nitems, int maxLen, int64_t outBase,
                        float* preac, float2* preconj, uint32_t* mask, int maskStride, cudaStream_t st)
{
    if (nitems <= 0 || maxLen <= 0) return;
    const int nseg = (maxLen + PSEG * 32 - 1) / (PSEG * 32);
    const int64_t warps = (int64_t)nitems * nseg;
    k_presiso<<<(unsigned)((warps + PWARPS - 1) / PWARPS), PWARPS * 32, 0, st>>>(iq, d_off, d_len, nitems, nseg, outBase, preac, preconj,
                                                                               mask, maskStride);
}

void c8b_launch_trigger(const float* preac, int64_t n, uint8_t* out, cudaStream_t st) { k_trigger<<<1, 32, 0, st>>>(preac, n, out); }

void c8b_launch_detect(const c8b_lut* lut, const float2* iq, const int64_t* d_off, const int32_t* d_len, int nitems, int itemBase,
                       int maxf, int64_t outBase, const float* preac, const uint32_t* mask, int maskStride, c8b_frame* frames,
                       float2* chan, cudaStream_t st)
{
    if (nitems <= 0) return;
    k_detect<<<(nitems + 63) / 64, 64, 0, st>>>(lut, iq, d_off, d_len, nitems, itemBase, maxf, outBase, preac, mask, maskStride, frames, chan);
}

void c8b_launch_header(const c8b_lut* lut, const float2* iq, const int64_t* d_off, int nitems, int maxf, int mupos, c8b_frame* frames,
                       const float2* chan, float2* hinv, int64_t llrStride, cudaStream_t st)
{
    if (nitems <= 0) return;
    k_header<<<(nitems * maxf + 63) / 64, 64, 0, st>>>(lut, iq, d_off, nitems * maxf, maxf, mupos, frames, chan, hinv, llrStride);
}
