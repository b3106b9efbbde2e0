// k_demod.cu -- the per-symbol hot loop of the demod block, fused with the signal block's CFO copy:
//   CFO rotation (lib/signal_impl.cc:164-192)  ->  64-point FFT of samples [8,72) of the symbol
//   (fftDemod, lib/demod_impl.cc:541-547)  ->  equalise with the frame's channel
//   (legacyChanUpdate :507-539 / nonLegacyChanUpdate :413-447)  ->  4-pilot common phase
//   ->  max-log LLR (procSymQamToLlr, lib/cloud80211phy.cc:2090-2148)  ->  deinterleave scatter
//   (procSymDeintL2 :2150-2192 / procSymDeintNL2SS1 :2238-2287)  ->  LLR stream.
// Symbols are independent (pilot polarity and pilot rotation are closed-form in the symbol index),
// so the grid is (symbol group, frame).  HBM traffic per symbol: 640 B of IQ in, 4*nCBPS B out.
//
// Mapping: 8 threads per symbol, 4 symbols per warp, 4 warps per CTA (16 consecutive symbols of one
// frame).  64 = 8 x 8: thread j takes samples j+8m, does an 8-point DFT in registers, multiplies by
// W64^(j*k1), the 8x8 transpose goes through padded (conflict-free) shared memory, and a second
// 8-point DFT leaves thread k1 with bins k1+8*k2.  LLRs are scattered through the deinterleave map into
// a shared-memory line per symbol and leave the SM as 16-byte coalesced stores.
#include "common.cuh"

namespace {

constexpr int DW = 4;               // warps per CTA
constexpr int SPW = 4;              // symbols per warp
constexpr int SPB = DW * SPW;       // symbols per CTA
constexpr int XS = 72;              // complex words per symbol in the transpose buffer (8 rows of 9)

struct cpx { float x, y; };
__device__ __forceinline__ cpx operator+(cpx a, cpx b) { return { a.x + b.x, a.y + b.y }; }
__device__ __forceinline__ cpx operator-(cpx a, cpx b) { return { a.x - b.x, a.y - b.y }; }
__device__ __forceinline__ cpx cmul(cpx a, cpx b) { return { a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x }; }
__device__ __forceinline__ cpx mulmj(cpx a) { return { a.y, -a.x }; }              // a * (-j)

// 8-point forward DFT, in place, natural order in and out
__device__ __forceinline__ void dft8(cpx* v)
{
    const float r = 0.70710678118654752440f;
    cpx a0 = v[0] + v[4], a1 = v[0] - v[4], a2 = v[2] + v[6], a3 = mulmj(v[2] - v[6]);
    cpx a4 = v[1] + v[5], a5 = v[1] - v[5], a6 = v[3] + v[7], a7 = mulmj(v[3] - v[7]);
    cpx b0 = a0 + a2, b2 = a0 - a2, b1 = a1 + a3, b3 = a1 - a3;
    cpx b4 = a4 + a6, b6 = mulmj(a4 - a6), b5 = a5 + a7, b7 = a5 - a7;
    b5 = { (b5.x + b5.y) * r, (b5.y - b5.x) * r };                                  // * (1-j)/sqrt2
    b7 = { (b7.y - b7.x) * r, -(b7.x + b7.y) * r };                                 // * (-1-j)/sqrt2
    v[0] = b0 + b4; v[4] = b0 - b4; v[1] = b1 + b5; v[5] = b1 - b5;
    v[2] = b2 + b6; v[6] = b2 - b6; v[3] = b3 + b7; v[7] = b3 - b7;
}

struct __align__(16) WarpBuf {
    float2 xch[SPW * XS];           // transpose buffer
    float2 pil[SPW * 4];            // equalised pilots of each symbol: bins 7, 21, 43, 57
    float llr[SPW * 416];           // deinterleaved soft bits of the warp's symbols, contiguous
};

__global__ void __launch_bounds__(DW * 32)
k_demod(const c8b_lut* __restrict__ lut, const float2* __restrict__ iq, const int64_t* __restrict__ off,
        const c8b_frame* __restrict__ frames, const float2* __restrict__ hinvAll, float* __restrict__ llrArena)
{
    __shared__ WarpBuf wb[DW];
    __shared__ uint16_t smap[416];
    const int item = blockIdx.y;
    const c8b_frame* __restrict__ fr = frames + item;
    if (fr->status != C8B_ST_OK) return;
    const int nsym = fr->nsym;
    const int sym0 = blockIdx.x * SPB;
    if (sym0 >= nsym) return;
    const int fmt = fr->format, ncbps = fr->ncbps, nss = fr->nss;
    const bool legacy = fmt == C8B_F_L;
    const int nbpsc = legacy ? ncbps / 48 : ncbps / (52 * (nss > 0 ? nss : 1));
    const int mi = nbpsc == 1 ? 0 : nbpsc == 2 ? 1 : nbpsc == 4 ? 2 : nbpsc == 6 ? 3 : 4;
    {   // deinterleave map of this frame
        const uint16_t* __restrict__ src = legacy ? lut->deintL[mi > 3 ? 3 : mi] : lut->deintNL[0][mi];
        for (int i = threadIdx.x; i < ncbps && i < 416; i += DW * 32) smap[i] = src[i];
    }
    __syncthreads();

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 3, j = lane & 7;                 // symbol within warp, thread within symbol
    WarpBuf& W = wb[warp];
    const int sidx = sym0 + warp * SPW + g;
    const bool live = sidx < nsym;
    const int nsymsamp = fr->nsymsamp;
    const float rad = fr->rad;
    const int k0 = fr->data_off + sidx * nsymsamp + C8B_SYM_SHIFT;       // index in the signal block's output stream
    const float2* __restrict__ x = iq + off[item] + fr->sync_idx + 224 + k0;

    cpx v[8];
    if (live) {
#pragma unroll
        for (int m = 0; m < 8; m++) {
            const int n = j + 8 * m;
            const float2 s = __ldg(x + n);
            float sn, cs;
            sincosf(__fmul_rn((float)(k0 + n + 224), rad), &sn, &cs);     // lib/signal_impl.cc:172-173
            v[m] = { s.x * cs - s.y * sn, s.x * sn + s.y * cs };
        }
    } else {
#pragma unroll
        for (int m = 0; m < 8; m++) v[m] = { 0.f, 0.f };
    }
    dft8(v);
#pragma unroll
    for (int k1 = 1; k1 < 8; k1++) {
        const int t = (j * k1) & 63;
        v[k1] = cmul(v[k1], cpx{ __ldg(&lut->twr[t]), __ldg(&lut->twi[t]) });
    }
#pragma unroll
    for (int k1 = 0; k1 < 8; k1++) W.xch[g * XS + k1 * 9 + j] = make_float2(v[k1].x, v[k1].y);
    __syncwarp();
#pragma unroll
    for (int n1 = 0; n1 < 8; n1++) { const float2 t = W.xch[g * XS + j * 9 + n1]; v[n1] = { t.x, t.y }; }
    dft8(v);                                               // v[k2] = bin j + 8*k2

    // equalise: s = F * (1/H)
    const float2* __restrict__ hinv = hinvAll + (size_t)item * 64;
#pragma unroll
    for (int k2 = 0; k2 < 8; k2++) {
        const float2 h = __ldg(hinv + j + 8 * k2);
        v[k2] = cmul(v[k2], cpx{ h.x, h.y });
    }
    // pilots: bin 7 = (j 7, k2 0), 21 = (5, 2), 43 = (3, 5), 57 = (1, 7)
    if (j == 7) W.pil[g * 4 + 0] = make_float2(v[0].x, v[0].y);
    if (j == 5) W.pil[g * 4 + 1] = make_float2(v[2].x, v[2].y);
    if (j == 3) W.pil[g * 4 + 2] = make_float2(v[5].x, v[5].y);
    if (j == 1) W.pil[g * 4 + 3] = make_float2(v[7].x, v[7].y);
    __syncwarp();
    cpx ps;
    {
        // pilot values: base {1,1,1,-1}; HT/VHT rotate left once per symbol (pilotShift, demod_impl.cc:549-557);
        // polarity index starts at 1 (L), 3 (HT), 4 (VHT) (demod_impl.cc:214,191,164)
        const int p0 = legacy ? 1 : (fmt == C8B_F_HT ? 3 : 4);
        const float P = __ldg(&lut->pilotP[(p0 + sidx) % 127]);
        const int sh = legacy ? 0 : (sidx & 3);
        // pilot[m] of this symbol = base[(m + sh) & 3], base[3] = -1
        const float q2 = (((2 + sh) & 3) == 3 ? -P : P), q3 = (((3 + sh) & 3) == 3 ? -P : P);
        const float q0 = (((0 + sh) & 3) == 3 ? -P : P), q1 = (((1 + sh) & 3) == 3 ? -P : P);
        const float2 s7 = W.pil[g * 4 + 0], s21 = W.pil[g * 4 + 1], s43 = W.pil[g * 4 + 2], s57 = W.pil[g * 4 + 3];
        float re = __fadd_rn(__fadd_rn(__fadd_rn(s7.x * q2, s21.x * q3), s43.x * q0), s57.x * q1);
        float im = __fadd_rn(__fadd_rn(__fadd_rn(s7.y * q2, s21.y * q3), s43.y * q0), s57.y * q1);
        const float inv = 1.0f / sqrtf(re * re + im * im);
        ps = { re * inv, -im * inv };                      // conj(sum) / |sum|
    }
    // soft bits of this thread's data tones, scattered through the deinterleave map
    const uint8_t* __restrict__ b2d = legacy ? lut->binToDataL : lut->binToDataNL;
    float* __restrict__ L = W.llr + g * ncbps;
    if (live) {
#pragma unroll
        for (int k2 = 0; k2 < 8; k2++) {
            const int d = b2d[j + 8 * k2];
            if (d == 255) continue;
            cpx q = cmul(v[k2], ps);
            const uint16_t* __restrict__ mp = smap + d * nbpsc;
            if (nbpsc == 1) {
                L[mp[0]] = q.x;
            } else if (nbpsc == 2) {
                q = { q.x * 1.4142135623730951f, q.y * 1.4142135623730951f };
                L[mp[0]] = q.x; L[mp[1]] = q.y;
            } else if (nbpsc == 4) {
                q = { q.x * 3.1622776601683795f, q.y * 3.1622776601683795f };
                L[mp[0]] = q.x; L[mp[1]] = 2.0f - fabsf(q.x);
                L[mp[2]] = q.y; L[mp[3]] = 2.0f - fabsf(q.y);
            } else if (nbpsc == 6) {
                q = { q.x * 6.48074069840786f, q.y * 6.48074069840786f };
                const float a = 4.0f - fabsf(q.x), b = 4.0f - fabsf(q.y);
                L[mp[0]] = q.x; L[mp[1]] = a; L[mp[2]] = 2.0f - fabsf(a);
                L[mp[3]] = q.y; L[mp[4]] = b; L[mp[5]] = 2.0f - fabsf(b);
            } else {
                q = { q.x * 13.038404810405298f, q.y * 13.038404810405298f };
                const float a = 8.0f - fabsf(q.x), b = 8.0f - fabsf(q.y);
                const float a2 = 4.0f - fabsf(a), b2 = 4.0f - fabsf(b);
                L[mp[0]] = q.x; L[mp[1]] = a; L[mp[2]] = a2; L[mp[3]] = 2.0f - fabsf(a2);
                L[mp[4]] = q.y; L[mp[5]] = b; L[mp[6]] = b2; L[mp[7]] = 2.0f - fabsf(b2);
            }
        }
    }
    __syncwarp();
    // coalesced copy-out of the warp's live symbols
    const int wsym0 = sym0 + warp * SPW;
    int nlive = nsym - wsym0;
    nlive = nlive < 0 ? 0 : (nlive > SPW ? SPW : nlive);
    const int nfl = nlive * ncbps;                          // multiple of 4 (48 | 52 divide by 4)
    float* __restrict__ out = llrArena + fr->llr_off + (int64_t)wsym0 * ncbps;
    if ((reinterpret_cast<uintptr_t>(out) & 15) == 0) {
        const float4* __restrict__ s4 = reinterpret_cast<const float4*>(W.llr);
        float4* __restrict__ o4 = reinterpret_cast<float4*>(out);
        for (int i = lane; i < nfl / 4; i += 32) o4[i] = s4[i];
    } else {
        for (int i = lane; i < nfl; i += 32) out[i] = W.llr[i];
    }
}

}  // namespace

void c8b_launch_demod(const c8b_lut* lut, const float2* iq, const int64_t* d_off, int nitems, int maxSym, const c8b_frame* frames,
                      const float2* hinv, float* llr, cudaStream_t st)
{
    if (nitems <= 0 || maxSym <= 0) return;
    for (int base = 0; base < nitems; base += 65535) {
        const int cnt = nitems - base < 65535 ? nitems - base : 65535;
        dim3 grid((maxSym + SPB - 1) / SPB, cnt);
        k_demod<<<grid, DW * 32, 0, st>>>(lut, iq, d_off + base, frames + base, hinv + (size_t)base * 64, llr);
    }
}
