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

constexpr int PTHREADS = 256;
constexpr int PR = 8;                         // consecutive samples per thread
constexpr int PS16 = PTHREADS * PR;           // s16 values per CTA (2048)
constexpr int PHALO = 64;                     // s16 values kept only as history (48 used, padded to a multiple of PR)
constexpr int PT = PS16 - PHALO;              // outputs per CTA (1984)
constexpr int PXN = PS16 + 32;                // x tile: the first s16 needs 31 samples of history (15 products + 16 delay)
static_assert(PT % 32 == 0 && PHALO % 32 == 0 && PHALO >= 48, "bitmap words must align with CTA tiles");

// x tile in shared memory with 2 pad slots after every 8 complex samples: a thread's 16-byte reads start 20 words apart,
// which spreads the 8 lanes of a quarter-warp over all 32 banks
__device__ __forceinline__ int xpos(int i) { return i + 2 * (i >> 3); }

struct f3 { float x, y, z; };
__device__ __forceinline__ f3 add3(f3 a, f3 b) { return { __fadd_rn(a.x, b.x), __fadd_rn(a.y, b.y), __fadd_rn(a.z, b.z) }; }

// The moving sums are evaluated as a fixed sliding tree so every output is a pure function of the 64 (80) samples
// before it, independent of tiling: s2[n]=v[n-1]+v[n]; s4[n]=s2[n-2]+s2[n]; s8; s16;
// sum48[n]=(s16[n-32]+s16[n-16])+s16[n]; sum64[n]=(s16[n-48]+s16[n-32])+(s16[n-16]+s16[n]).
// A thread forms the 23 products its 8 consecutive s16 need and runs the tree in registers (no barrier per level);
// only s16 goes through shared memory for the 16/32/48-sample lags.
__global__ void __launch_bounds__(PTHREADS)
k_presiso(const float2* __restrict__ iq, const int64_t* __restrict__ off, const int32_t* __restrict__ len, int64_t outBase,
          float* __restrict__ preac, float2* __restrict__ preconj, uint32_t* __restrict__ mask, int maskStride)
{
    __shared__ __align__(16) float2 xs[PXN + 2 * (PXN / 8) + 8];
    __shared__ __align__(16) float s16re[PS16], s16im[PS16], s16pw[PS16];
    const int item = blockIdx.y;
    const int n = len[item];
    const int t0 = blockIdx.x * PT;                 // first output of this CTA
    if (t0 >= n) return;
    const float2* __restrict__ x = iq + off[item];
    const int64_t ob = off[item] - outBase;
    const int s0 = t0 - PHALO;                      // item index of s16 slot 0
    const int x0 = s0 - 32;                         // item index of x tile slot 0 (multiple of 8 relative to s0)
    for (int j = threadIdx.x; j < PXN; j += PTHREADS) {
        const int i = x0 + j;
        xs[xpos(j)] = (i >= 0 && i < n) ? x[i] : make_float2(0.f, 0.f);
    }
    __syncthreads();
    const int q0 = threadIdx.x * PR;                // this thread's s16 slots q0 .. q0+7  <->  item index s0 + q0 + k
    {
        // samples x[s0+q0-31 .. s0+q0+7] = tile slots q0+1 .. q0+39; load slots q0 .. q0+39 as 20 aligned pairs
        float2 xv[40];
#pragma unroll
        for (int k = 0; k < 40; k += 2) {
            const float4 t = *reinterpret_cast<const float4*>(&xs[xpos(q0 + k)]);
            xv[k] = make_float2(t.x, t.y); xv[k + 1] = make_float2(t.z, t.w);
        }
        f3 v[23];                                   // v[m] <-> item index s0 + q0 - 15 + m  (tile slot q0 + 17 + m)
#pragma unroll
        for (int m = 0; m < 23; m++) {
            const float2 c = xv[17 + m], d = xv[1 + m];     // x[i], x[i-16]   (delay(16), in0 * conj(in1))
            v[m].x = __fadd_rn(__fmul_rn(d.x, c.x), __fmul_rn(d.y, c.y));
            v[m].y = __fsub_rn(__fmul_rn(d.y, c.x), __fmul_rn(d.x, c.y));
            v[m].z = __fadd_rn(__fmul_rn(c.x, c.x), __fmul_rn(c.y, c.y));       // complex_to_mag_squared
        }
#pragma unroll
        for (int m = 22; m >= 1; m--) v[m] = add3(v[m - 1], v[m]);              // s2 at m >= 1
#pragma unroll
        for (int m = 22; m >= 3; m--) v[m] = add3(v[m - 2], v[m]);              // s4 at m >= 3
#pragma unroll
        for (int m = 22; m >= 7; m--) v[m] = add3(v[m - 4], v[m]);              // s8 at m >= 7
#pragma unroll
        for (int m = 22; m >= 15; m--) v[m] = add3(v[m - 8], v[m]);             // s16 at m >= 15  <->  slots q0 .. q0+7
        *reinterpret_cast<float4*>(&s16re[q0]) = make_float4(v[15].x, v[16].x, v[17].x, v[18].x);
        *reinterpret_cast<float4*>(&s16re[q0 + 4]) = make_float4(v[19].x, v[20].x, v[21].x, v[22].x);
        *reinterpret_cast<float4*>(&s16im[q0]) = make_float4(v[15].y, v[16].y, v[17].y, v[18].y);
        *reinterpret_cast<float4*>(&s16im[q0 + 4]) = make_float4(v[19].y, v[20].y, v[21].y, v[22].y);
        *reinterpret_cast<float4*>(&s16pw[q0]) = make_float4(v[15].z, v[16].z, v[17].z, v[18].z);
        *reinterpret_cast<float4*>(&s16pw[q0 + 4]) = make_float4(v[19].z, v[20].z, v[21].z, v[22].z);
    }
    __syncthreads();
    const int i0 = s0 + q0;                         // first output index of this thread (multiple of 8)
    const bool outp = q0 >= PHALO && i0 < n;        // history-only slots and slots past the item produce nothing
    uint32_t bits = 0;
    if (outp) {
        float ac[PR];
        // 16-byte reads of the lagged s16 values (scalar reads at an 8-word thread stride would be 8-way bank conflicts)
        float re0[PR], re1[PR], re2[PR], im0[PR], im1[PR], im2[PR], pw0[PR], pw1[PR], pw2[PR], pw3[PR];
        auto ld8 = [&](const float* base, int q, float* dst) {
            const float4 a = *reinterpret_cast<const float4*>(base + q), b = *reinterpret_cast<const float4*>(base + q + 4);
            dst[0] = a.x; dst[1] = a.y; dst[2] = a.z; dst[3] = a.w; dst[4] = b.x; dst[5] = b.y; dst[6] = b.z; dst[7] = b.w;
        };
        ld8(s16re, q0 - 32, re2); ld8(s16re, q0 - 16, re1); ld8(s16re, q0, re0);
        ld8(s16im, q0 - 32, im2); ld8(s16im, q0 - 16, im1); ld8(s16im, q0, im0);
        ld8(s16pw, q0 - 48, pw3); ld8(s16pw, q0 - 32, pw2); ld8(s16pw, q0 - 16, pw1); ld8(s16pw, q0, pw0);
#pragma unroll
        for (int k = 0; k < PR; k++) {
            const float cr = __fadd_rn(__fadd_rn(re2[k], re1[k]), re0[k]);
            const float ci = __fadd_rn(__fadd_rn(im2[k], im1[k]), im0[k]);
            const float pw = __fadd_rn(__fadd_rn(pw3[k], pw2[k]), __fadd_rn(pw1[k], pw0[k]));
            const float mag = __fsqrt_rn(__fadd_rn(__fmul_rn(cr, cr), __fmul_rn(ci, ci)));   // complex_to_mag
            ac[k] = __fdiv_rn(mag, pw);                                                         // divide_ff
            if (ac[k] > 0.3f && i0 + k < n) bits |= 1u << k;                                    // lib/trigger_impl.cc:79
            if (preconj && i0 + k < n) preconj[ob + i0 + k] = make_float2(cr, ci);
        }
        if (i0 + PR <= n && ((ob + i0) & 3) == 0) {
            *reinterpret_cast<float4*>(&preac[ob + i0]) = make_float4(ac[0], ac[1], ac[2], ac[3]);
            *reinterpret_cast<float4*>(&preac[ob + i0 + 4]) = make_float4(ac[4], ac[5], ac[6], ac[7]);
        } else {
#pragma unroll
            for (int k = 0; k < PR; k++) if (i0 + k < n) preac[ob + i0 + k] = ac[k];
        }
    }
    if (mask) {
        // threshold bitmap for the trigger scan: s0 is a multiple of 32 (PT and PHALO are), so 4 neighbouring lanes
        // hold the 4 bytes of one word
        uint32_t w = bits << (8 * (threadIdx.x & 3));
        w |= __shfl_xor_sync(0xffffffffu, w, 1);
        w |= __shfl_xor_sync(0xffffffffu, w, 2);
        if ((threadIdx.x & 3) == 0 && outp) mask[(size_t)item * maskStride + (i0 >> 5)] = w;
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
    for (int base = 0; base < nitems; base += 65535) {
        const int cnt = nitems - base < 65535 ? nitems - base : 65535;
        dim3 grid((maxLen + PT - 1) / PT, cnt);
        k_presiso<<<grid, PTHREADS, 0, st>>>(iq, d_off + base, d_len + base, outBase, preac, preconj,
                                              mask ? mask + (size_t)base * maskStride : nullptr, maskStride);
    }
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
