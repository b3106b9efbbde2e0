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

constexpr int PT = 1024;           // outputs per CTA
constexpr int PH = 64;             // history of the 64-sample power window
constexpr int PTHREADS = 256;

// The moving sums are evaluated as a fixed sliding tree so every output is a pure function of the 64
// (80) samples before it, independent of tiling: s2[n]=v[n-1]+v[n]; s4[n]=s2[n-2]+s2[n]; s8; s16;
// sum48[n]=(s16[n-32]+s16[n-16])+s16[n]; sum64[n]=(s16[n-48]+s16[n-32])+(s16[n-16]+s16[n]).
__global__ void __launch_bounds__(PTHREADS)
k_presiso(const float2* __restrict__ iq, const int64_t* __restrict__ off, const int32_t* __restrict__ len, int64_t outBase,
          float* __restrict__ preac, float2* __restrict__ preconj, uint32_t* __restrict__ mask, int maskStride)
{
    __shared__ float4 s[PT + PH];                   // (re, im, |x|^2, -) of v at item index t0 - 64 + j
    const int item = blockIdx.y;
    const int n = len[item];
    const int t0 = blockIdx.x * PT;
    if (t0 >= n) return;
    const float2* __restrict__ x = iq + off[item];
    const int64_t ob = off[item] - outBase;
    // products for j in [0, PT+64): index i = t0 - 64 + j
    for (int j = threadIdx.x; j < PT + PH; j += PTHREADS) {
        const int i = t0 - PH + j;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (i >= 0 && i < n) {
            const float2 c = x[i];
            v.z = __fadd_rn(__fmul_rn(c.x, c.x), __fmul_rn(c.y, c.y));          // complex_to_mag_squared
            if (i >= 16) {
                const float2 d = x[i - 16];                                     // delay(16), in0 * conj(in1)
                v.x = __fadd_rn(__fmul_rn(d.x, c.x), __fmul_rn(d.y, c.y));
                v.y = __fsub_rn(__fmul_rn(d.y, c.x), __fmul_rn(d.x, c.y));
            }
        }
        s[j] = v;
    }
    __syncthreads();
    constexpr int PER = (PT + PH + PTHREADS - 1) / PTHREADS;
#pragma unroll
    for (int lag = 1; lag < 16; lag <<= 1) {
        float4 r[PER];
#pragma unroll
        for (int q = 0; q < PER; q++) {
            const int j = threadIdx.x + q * PTHREADS;
            if (j < PT + PH) {
                const float4 b = s[j];
                float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
                if (j >= lag) a = s[j - lag];
                r[q] = make_float4(__fadd_rn(a.x, b.x), __fadd_rn(a.y, b.y), __fadd_rn(a.z, b.z), 0.f);
            }
        }
        __syncthreads();
#pragma unroll
        for (int q = 0; q < PER; q++) {
            const int j = threadIdx.x + q * PTHREADS;
            if (j < PT + PH) s[j] = r[q];
        }
        __syncthreads();
    }
    // s[j] = s16 at index t0-64+j (exact for j >= 15; outputs use j >= 16)
    for (int k = threadIdx.x; k < PT; k += PTHREADS) {
        const int i = t0 + k;
        if (i >= n) break;
        const int j = k + PH;
        const float4 a0 = s[j], a1 = s[j - 16], a2 = s[j - 32], a3 = s[j - 48];
        const float cr = __fadd_rn(__fadd_rn(a2.x, a1.x), a0.x), ci = __fadd_rn(__fadd_rn(a2.y, a1.y), a0.y);
        const float pw = __fadd_rn(__fadd_rn(a3.z, a2.z), __fadd_rn(a1.z, a0.z));
        const float mag = __fsqrt_rn(__fadd_rn(__fmul_rn(cr, cr), __fmul_rn(ci, ci)));   // complex_to_mag
        const float ac = __fdiv_rn(mag, pw);                                                // divide_ff
        preac[ob + i] = ac;
        if (preconj) preconj[ob + i] = make_float2(cr, ci);
        if (mask) {                                        // threshold bitmap for the trigger scan (lib/trigger_impl.cc:79)
            // i is a multiple of 32 at lane 0 (t0 and k are); lanes past the end of the item left the loop above
            const uint32_t m = __ballot_sync(__activemask(), ac > 0.3f);
            if ((threadIdx.x & 31) == 0) mask[(size_t)item * maskStride + (i >> 5)] = m;
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
