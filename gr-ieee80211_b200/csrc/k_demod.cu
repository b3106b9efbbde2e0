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
// 8-point DFT leaves thread k1 with bins k1+8*k2.  LLRs are scattered through the deinterleaver -- in closed form: a
// per-thread table entry gives each data tone's base position and rotation, no map is staged or looked up -- into
// a shared-memory line per symbol and leave the SM as one bulk copy per warp (cp.async.bulk, see bulk_store).
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
    union {
        float2 xch[SPW * XS];       // transpose buffer (dead once the second DFT has read it)
        float llr[SPW * 416];       // deinterleaved soft bits of the warp's symbols, contiguous
    };
    float2 pil[SPW * 4];            // equalised pilots of each symbol: bins 7, 21, 43, 57
};

// cos / sin of the CFO phase the reference evaluates with cosf / sinf (lib/signal_impl.cc:172-173).  The phase itself is the
// reference's float product (up to ~150 rad, so its own rounding is ~1e-5 rad); reduced to [-pi, pi] with a two-constant
// Cody-Waite step it goes through the SFU approximations, whose 4e-7 absolute error is far inside the 1e-4 LLR tolerance
// and an order of magnitude cheaper than the full-range sincosf.
__device__ __forceinline__ void cfo_rot(float ph, float* sn, float* cs)
{
    const float k = rintf(ph * 0.15915494309189535f);
    float r = fmaf(-k, 6.2831854820251465f, ph);            // 2 pi = 6.2831854820251465 - 1.7484555e-7
    r = fmaf(k, 1.7484555e-7f, r);
    *sn = __sinf(r);
    *cs = __cosf(r);
}

// shared memory -> global as ONE bulk copy by the copy engine (cp.async.bulk, SASS UBLKCP): the line the warp scattered leaves
// the SM without passing the load / store pipe again (the lane loop costs an LDS.128 + STG.128 wavefront per 128 bytes, a
// quarter of this kernel's L1 traffic).  src / dst 16-byte aligned, bytes a multiple of 16.  Every lane that wrote the line has
// fenced its writes towards the async proxy and the warp has synchronised before one lane calls this.
__device__ __forceinline__ void bulk_store(float* dst, const float* src, uint32_t bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                 :: "l"(dst), "r"((uint32_t)__cvta_generic_to_shared(src)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_store_fence() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void bulk_store_drain()            // the shared-memory side of every bulk copy of this thread is done
{
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}

// soft bits of one data tone through the deinterleaver in closed form (lut.h demapTab): NB bits per tone, s = max(NB/2, 1) per
// axis, axis h starts NCOL*s floats after axis 0, bit c of an axis sits at NCOL*((c + R) mod s).  NCOL is 16 (legacy) or 13.
template <int NCOL, int NB>
__device__ __forceinline__ void demap_tone(float* __restrict__ Lb, int R, cpx q)
{
    if (NB == 1) {
        Lb[0] = q.x;
    } else if (NB == 2) {
        Lb[0] = q.x * 1.4142135623730951f; Lb[NCOL] = q.y * 1.4142135623730951f;
    } else if (NB == 4 && NCOL == 13) {
        // HT/VHT 16 / 64-QAM: the axis' VALUES are rotated by R and stored at the fixed offsets NCOL*m.  With the offsets the
        // same in every lane the stores of a warp spread over the banks (bank model: 64-QAM 1.25 wavefronts per store against
        // 1.75 with rotated addresses; measured 1.036 -> 1.027 ms per chunk); the other maps gain nothing from it.
        q = { q.x * 3.1622776601683795f, q.y * 3.1622776601683795f };
        const float a = 2.0f - fabsf(q.x), b = 2.0f - fabsf(q.y);
        const bool r1 = R != 0;
        Lb[0] = r1 ? a : q.x; Lb[NCOL] = r1 ? q.x : a;
        Lb[2 * NCOL] = r1 ? b : q.y; Lb[3 * NCOL] = r1 ? q.y : b;
    } else if (NB == 6 && NCOL == 13) {
        q = { q.x * 6.48074069840786f, q.y * 6.48074069840786f };
        const float a = 4.0f - fabsf(q.x), b = 4.0f - fabsf(q.y);
        const float a2 = 2.0f - fabsf(a), b2 = 2.0f - fabsf(b);
        const bool r0 = R == 0, r1 = R == 1;                               // position m holds value (m - R) mod 3
        Lb[0] = r0 ? q.x : (r1 ? a2 : a);
        Lb[NCOL] = r0 ? a : (r1 ? q.x : a2);
        Lb[2 * NCOL] = r0 ? a2 : (r1 ? a : q.x);
        Lb[3 * NCOL] = r0 ? q.y : (r1 ? b2 : b);
        Lb[4 * NCOL] = r0 ? b : (r1 ? q.y : b2);
        Lb[5 * NCOL] = r0 ? b2 : (r1 ? b : q.y);
    } else if (NB == 4) {                                                  // s = 2: R in {0, 1}
        q = { q.x * 3.1622776601683795f, q.y * 3.1622776601683795f };
        float* __restrict__ A0 = Lb + R * NCOL;
        float* __restrict__ A1 = Lb + (R ^ 1) * NCOL;
        A0[0] = q.x; A1[0] = 2.0f - fabsf(q.x);
        A0[2 * NCOL] = q.y; A1[2 * NCOL] = 2.0f - fabsf(q.y);
    } else if (NB == 6) {                                                  // s = 3: R in {0, 1, 2}
        q = { q.x * 6.48074069840786f, q.y * 6.48074069840786f };
        const float a = 4.0f - fabsf(q.x), b = 4.0f - fabsf(q.y);
        const int t1 = R == 2 ? 0 : R + 1, t2 = R == 0 ? 2 : R - 1;
        float* __restrict__ A0 = Lb + R * NCOL;
        float* __restrict__ A1 = Lb + t1 * NCOL;
        float* __restrict__ A2 = Lb + t2 * NCOL;
        A0[0] = q.x; A1[0] = a; A2[0] = 2.0f - fabsf(a);
        A0[3 * NCOL] = q.y; A1[3 * NCOL] = b; A2[3 * NCOL] = 2.0f - fabsf(b);
    } else {                                                               // s = 4: R in {0 .. 3}
        q = { q.x * 13.038404810405298f, q.y * 13.038404810405298f };
        const float a = 8.0f - fabsf(q.x), b = 8.0f - fabsf(q.y);
        const float a2 = 4.0f - fabsf(a), b2 = 4.0f - fabsf(b);
        float* __restrict__ A0 = Lb + R * NCOL;
        float* __restrict__ A1 = Lb + ((R + 1) & 3) * NCOL;
        float* __restrict__ A2 = Lb + ((R + 2) & 3) * NCOL;
        float* __restrict__ A3 = Lb + ((R + 3) & 3) * NCOL;
        A0[0] = q.x; A1[0] = a; A2[0] = a2; A3[0] = 2.0f - fabsf(a2);
        A0[4 * NCOL] = q.y; A1[4 * NCOL] = b; A2[4 * NCOL] = b2; A3[4 * NCOL] = 2.0f - fabsf(b2);
    }
}

template <int NCOL, int NB>
__device__ __forceinline__ void demap_symbol(float* __restrict__ L, const uint4 te, const cpx* v, cpx ps)
{
    const unsigned w[4] = { te.x, te.y, te.z, te.w };
#pragma unroll
    for (int k2 = 0; k2 < 8; k2++) {
        const unsigned e = (k2 & 1) ? (w[k2 >> 1] >> 16) : (w[k2 >> 1] & 0xFFFFu);
        if (e == 0xFFFFu) continue;                                        // null / pilot bin
        demap_tone<NCOL, NB>(L + (e & 511u), (int)(e >> 9), cmul(v[k2], ps));
    }
}

__global__ void __launch_bounds__(DW * 32, 8)
k_demod(const c8b_lut* __restrict__ lut, const float2* __restrict__ iq, const int64_t* __restrict__ off, int maxf,
        const c8b_frame* __restrict__ frames, const float2* __restrict__ hinvAll, float* __restrict__ llrArena)
{
    __shared__ WarpBuf wb[DW];
    const int item = blockIdx.y;
    const c8b_frame* __restrict__ fr = frames + item;
    // every word whose address follows from the block index is requested up front: one round trip instead of a chain of them
    const int status = fr->status, nss = fr->nss, nsym = fr->nsym, fmt = fr->format, ncbps = fr->ncbps;
    const int nsymsamp = fr->nsymsamp, data_off = fr->data_off, sync_idx = fr->sync_idx;
    const float rad = fr->rad;
    const int64_t llr_off = fr->llr_off;
    const int64_t ioff = off[maxf == 1 ? item : item / maxf];
    if (status != C8B_ST_OK || nss != 1) return;          // 2-stream frames: k_demod2
    const int sym0 = blockIdx.x * SPB;
    if (sym0 >= nsym) return;
    const bool legacy = fmt == C8B_F_L;
    const int nbpsc = legacy ? ncbps / 48 : ncbps / 52;
    const int mi = nbpsc == 1 ? 0 : nbpsc == 2 ? 1 : nbpsc == 4 ? 2 : nbpsc == 6 ? 3 : 4;

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 3, j = lane & 7;                 // symbol within warp, thread within symbol
    WarpBuf& W = wb[warp];
    const int sidx = sym0 + warp * SPW + g;
    const bool live = sidx < nsym;
    const int k0 = data_off + sidx * nsymsamp + C8B_SYM_SHIFT;           // index in the signal block's output stream
    const float2* __restrict__ x = iq + ioff + sync_idx + 224 + k0;      // blockIdx.y = frame slot

    cpx v[8];
    if (live) {
        float2 s[8];
#pragma unroll
        for (int m = 0; m < 8; m++) s[m] = __ldg(x + j + 8 * m);
#pragma unroll
        for (int m = 0; m < 8; m++) {
            float sn, cs;
            cfo_rot(__fmul_rn((float)(k0 + j + 8 * m + 224), rad), &sn, &cs);     // lib/signal_impl.cc:172-173
            v[m] = { s[m].x * cs - s[m].y * sn, s[m].x * sn + s[m].y * cs };
        }
    } else {
#pragma unroll
        for (int m = 0; m < 8; m++) v[m] = { 0.f, 0.f };
    }
    dft8(v);
    {
        const float4* __restrict__ tw = reinterpret_cast<const float4*>(lut->tw8[0][j]);      // W64^(j*k1), k1 = 2p, 2p + 1
#pragma unroll
        for (int p = 0; p < 4; p++) {
            const float4 t = __ldg(tw + 8 * p);
            if (p) v[2 * p] = cmul(v[2 * p], cpx{ t.x, t.y });
            v[2 * p + 1] = cmul(v[2 * p + 1], cpx{ t.z, t.w });
        }
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
    // soft bits of this thread's data tones, scattered through the deinterleaver (closed form: demapTab) into the symbol's
    // shared-memory row: ncbps floats + LPAD.  With rows at multiples of 48 * nbpsc the four symbols of a warp scatter
    // through the legacy deinterleaver into the same banks (11.5 wavefronts per store, 5.75 for BPSK; HT/VHT BPSK 2.1); four
    // floats of padding bring that to 2.9 / 1.9 / 1.4 (bank model over the lane -> address map; the other HT/VHT maps are at
    // 1.1-1.75 unpadded, and the 256-QAM row has no room to spare).  Multiples of 4 keep the rows float4-aligned.
    const int lpad = (legacy || nbpsc == 1) ? 4 : 0;
    const int lrow = ncbps + lpad;
    float* __restrict__ L = W.llr + g * lrow;
    if (live) {
        const uint4 te = __ldg(reinterpret_cast<const uint4*>(lut->demapTab[(legacy ? 0 : 4) + mi] + 8 * j));
        if (legacy) {
            if (nbpsc == 1) demap_symbol<16, 1>(L, te, v, ps);
            else if (nbpsc == 2) demap_symbol<16, 2>(L, te, v, ps);
            else if (nbpsc == 4) demap_symbol<16, 4>(L, te, v, ps);
            else demap_symbol<16, 6>(L, te, v, ps);
        } else {
            if (nbpsc == 1) demap_symbol<13, 1>(L, te, v, ps);
            else if (nbpsc == 2) demap_symbol<13, 2>(L, te, v, ps);
            else if (nbpsc == 4) demap_symbol<13, 4>(L, te, v, ps);
            else if (nbpsc == 6) demap_symbol<13, 6>(L, te, v, ps);
            else demap_symbol<13, 8>(L, te, v, ps);
        }
    }
    bulk_store_fence();
    __syncwarp();
    // copy-out of the warp's live symbols: one bulk copy (per row when the rows are padded), lane loop when the frame's soft-bit
    // stream is not 16-byte aligned in the arena
    const int wsym0 = sym0 + warp * SPW;
    int nlive = nsym - wsym0;
    nlive = nlive < 0 ? 0 : (nlive > SPW ? SPW : nlive);
    const int nfl = nlive * ncbps;                          // multiple of 4 (48 | 52 divide by 4)
    float* __restrict__ out = llrArena + llr_off + (int64_t)wsym0 * ncbps;
    const bool al16 = (reinterpret_cast<uintptr_t>(out) & 15) == 0;
    if (al16) {
        if (lane == 0 && nlive > 0) {
            if (lpad == 0) bulk_store(out, W.llr, (uint32_t)nfl * 4u);
            else for (int r = 0; r < nlive; r++) bulk_store(out + r * ncbps, W.llr + r * lrow, (uint32_t)ncbps * 4u);
            bulk_store_drain();
        }
    } else if (lpad == 0) {
        for (int i = lane; i < nfl; i += 32) out[i] = W.llr[i];
    } else {
        for (int r = 0; r < nlive; r++)
            for (int i = lane; i < ncbps; i += 32) out[r * ncbps + i] = W.llr[r * lrow + i];
    }
}

// ---------------------------------------------------------------------------------------------------
// k_demod2: per-symbol loop of the 2x2 block (lib/demod2_impl.cc:279-330; htChanUpdate :471-551,
// vhtChanUpdate :553-630; procSymDeintNL2SS1/SS2 c8p.cc:2238-2338; procSymDepasNL :2442-2451) for
// 2-stream frames.  16 threads per symbol: threads 0-7 transform antenna 0, threads 8-15 antenna 1
// (same 8x8 DFT as k_demod); the two halves swap spectra with 8 shuffles, then half a computes
// stream a = F0*w[2a] + F1*w[2a+1] (the folded (H^H H)^-1 H^H), the 8-pilot common phase is shared
// through shared memory, and each half demaps / deinterleaves its own stream straight into the
// stream-deparsed position of the symbol's LLR line (closed form: demap2_tone).
// HBM traffic per symbol: 2 x 640 B in, 4*nCBPS B out (3776 B at HT MCS15).
// ---------------------------------------------------------------------------------------------------
constexpr int SPW2 = 2;             // symbols per warp
constexpr int SPB2 = DW * SPW2;     // symbols per CTA

struct __align__(16) WarpBuf2 {
    float2 xch[SPW2 * 2 * XS];      // transpose buffers: [symbol][antenna]
    float2 pil[SPW2 * 8];           // equalised pilots: [symbol][stream][7, 21, 43, 57]
    float llr[SPW2 * 832];          // deparsed soft bits of the warp's symbols, contiguous
};

// One data tone of one stream of a 2-stream symbol: NB soft bits through the stream's deinterleaver (c8p.cc:2238-2338) and the
// stream parser (c8p.cc:2442-2451) in closed form.  Stream-local position of bit h*S + c: k = B + 13*((c + R) mod S) + 13*S*h
// (as in demap_tone); the parser sends k to k + S*floor(k / S) + a*S.  The table entry carries P0 = the parsed position of k = B,
// R and beta = B mod S, so position m of axis h is P0 + 13*m + S*floor((beta + 13*m) / S) + 26*S*h and holds value (m - R) mod S.
template <int NB>
__device__ __forceinline__ void demap2_tone(float* __restrict__ Lb, int R, int beta, cpx q)
{
    constexpr int S = NB / 2 > 1 ? NB / 2 : 1;
    if (NB == 1) {
        Lb[0] = q.x;
    } else if (NB == 2) {
        Lb[0] = q.x * 1.4142135623730951f; Lb[26] = q.y * 1.4142135623730951f;
    } else if (NB == 4) {
        q = { q.x * 3.1622776601683795f, q.y * 3.1622776601683795f };
        const float a = 2.0f - fabsf(q.x), b = 2.0f - fabsf(q.y);
        const bool r1 = R != 0;
        float* __restrict__ A1 = Lb + 13 + S * ((beta + 13) / S);
        Lb[0] = r1 ? a : q.x; A1[0] = r1 ? q.x : a;
        Lb[26 * S] = r1 ? b : q.y; A1[26 * S] = r1 ? q.y : b;
    } else if (NB == 6) {
        q = { q.x * 6.48074069840786f, q.y * 6.48074069840786f };
        const float a = 4.0f - fabsf(q.x), b = 4.0f - fabsf(q.y);
        const float a2 = 2.0f - fabsf(a), b2 = 2.0f - fabsf(b);
        const bool r0 = R == 0, r1 = R == 1;
        float* __restrict__ A1 = Lb + 13 + S * ((beta + 13) / S);
        float* __restrict__ A2 = Lb + 26 + S * ((beta + 26) / S);
        Lb[0] = r0 ? q.x : (r1 ? a2 : a);
        A1[0] = r0 ? a : (r1 ? q.x : a2);
        A2[0] = r0 ? a2 : (r1 ? a : q.x);
        Lb[26 * S] = r0 ? q.y : (r1 ? b2 : b);
        A1[26 * S] = r0 ? b : (r1 ? q.y : b2);
        A2[26 * S] = r0 ? b2 : (r1 ? b : q.y);
    } else {
        q = { q.x * 13.038404810405298f, q.y * 13.038404810405298f };
        const float a = 8.0f - fabsf(q.x), b = 8.0f - fabsf(q.y);
        const float a2 = 4.0f - fabsf(a), b2 = 4.0f - fabsf(b);
        const float a3 = 2.0f - fabsf(a2), b3 = 2.0f - fabsf(b2);
        // value c goes to position (c + R) & 3
        const int m0 = R, m1 = (R + 1) & 3, m2 = (R + 2) & 3, m3 = (R + 3) & 3;
        float* __restrict__ A0 = Lb + 13 * m0 + S * ((beta + 13 * m0) >> 2);
        float* __restrict__ A1 = Lb + 13 * m1 + S * ((beta + 13 * m1) >> 2);
        float* __restrict__ A2 = Lb + 13 * m2 + S * ((beta + 13 * m2) >> 2);
        float* __restrict__ A3 = Lb + 13 * m3 + S * ((beta + 13 * m3) >> 2);
        A0[0] = q.x; A1[0] = a; A2[0] = a2; A3[0] = a3;
        A0[26 * S] = q.y; A1[26 * S] = b; A2[26 * S] = b2; A3[26 * S] = b3;
    }
}

template <int NB>
__device__ __forceinline__ void demap2_symbol(float* __restrict__ L, const uint4 te, const cpx* v, cpx ps)
{
    const unsigned w[4] = { te.x, te.y, te.z, te.w };
#pragma unroll
    for (int k2 = 0; k2 < 8; k2++) {
        const unsigned e = (k2 & 1) ? (w[k2 >> 1] >> 16) : (w[k2 >> 1] & 0xFFFFu);
        if (e == 0xFFFFu) continue;                                        // null / pilot bin
        demap2_tone<NB>(L + (e & 1023u), (int)((e >> 10) & 3u), (int)(e >> 12), cmul(v[k2], ps));
    }
}

__global__ void __launch_bounds__(DW * 32)
k_demod2(const c8b_lut* __restrict__ lut, const float2* __restrict__ iq0, const float2* __restrict__ iq1,
         const int64_t* __restrict__ off, int maxf, const c8b_frame* __restrict__ frames, const float2* __restrict__ w2All,
         float* __restrict__ llrArena)
{
    __shared__ WarpBuf2 wb[DW];
    const int item = blockIdx.y;
    const c8b_frame* __restrict__ fr = frames + item;
    // every word whose address follows from the block index is requested up front (one round trip, not a chain)
    const int status = fr->status, nss = fr->nss, nsym = fr->nsym, fmt = fr->format, ncbps = fr->ncbps;
    const int nsymsamp = fr->nsymsamp, data_off = fr->data_off, sync_idx = fr->sync_idx;
    const float rad = fr->rad;
    const int64_t llr_off = fr->llr_off;
    const int64_t ioff = off[maxf == 1 ? item : item / maxf];
    if (status != C8B_ST_OK || nss != 2) return;
    const int sym0 = blockIdx.x * SPB2;
    if (sym0 >= nsym) return;
    const int ncbpss = ncbps >> 1;
    const int nbpsc = ncbpss / 52;
    const int mi = nbpsc == 1 ? 0 : nbpsc == 2 ? 1 : nbpsc == 4 ? 2 : nbpsc == 6 ? 3 : 4;

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 4, a = (lane >> 3) & 1, j = lane & 7;   // symbol in warp, antenna / stream, thread in the DFT
    WarpBuf2& W = wb[warp];
    const int sidx = sym0 + warp * SPW2 + g;
    const bool live = sidx < nsym;
    const int k0 = data_off + sidx * nsymsamp + C8B_SYM_SHIFT;
    const float2* __restrict__ x = (a ? iq1 : iq0) + ioff + sync_idx + 224 + k0;

    cpx v[8];
    if (live) {
        float2 s[8];
#pragma unroll
        for (int m = 0; m < 8; m++) s[m] = __ldg(x + j + 8 * m);
#pragma unroll
        for (int m = 0; m < 8; m++) {
            float sn, cs;
            cfo_rot(__fmul_rn((float)(k0 + j + 8 * m + 224), rad), &sn, &cs);     // lib/signal2_impl.cc:172-177
            v[m] = { s[m].x * cs - s[m].y * sn, s[m].x * sn + s[m].y * cs };
        }
    } else {
#pragma unroll
        for (int m = 0; m < 8; m++) v[m] = { 0.f, 0.f };
    }
    dft8(v);
    {
        const float4* __restrict__ tw = reinterpret_cast<const float4*>(lut->tw8[0][j]);      // W64^(j*k1), k1 = 2p, 2p + 1
#pragma unroll
        for (int p = 0; p < 4; p++) {
            const float4 t = __ldg(tw + 8 * p);
            if (p) v[2 * p] = cmul(v[2 * p], cpx{ t.x, t.y });
            v[2 * p + 1] = cmul(v[2 * p + 1], cpx{ t.z, t.w });
        }
    }
    float2* __restrict__ xb = W.xch + (g * 2 + a) * XS;
#pragma unroll
    for (int k1 = 0; k1 < 8; k1++) xb[k1 * 9 + j] = make_float2(v[k1].x, v[k1].y);
    __syncwarp();
#pragma unroll
    for (int n1 = 0; n1 < 8; n1++) { const float2 t = xb[j * 9 + n1]; v[n1] = { t.x, t.y }; }
    dft8(v);                                               // v[k2] = bin j + 8*k2 of antenna a

    // zero-forcing: stream a = F_ant0 * w[2a] + F_ant1 * w[2a+1]
    const float2* __restrict__ w2 = w2All + (size_t)item * 264;
#pragma unroll
    for (int k2 = 0; k2 < 8; k2++) {
        cpx o;
        o.x = __shfl_xor_sync(0xffffffffu, v[k2].x, 8);
        o.y = __shfl_xor_sync(0xffffffffu, v[k2].y, 8);
        const cpx f0 = a ? o : v[k2], f1 = a ? v[k2] : o;
        const float2 wa = __ldg(w2 + 4 * (j + 8 * k2) + 2 * a), wb2 = __ldg(w2 + 4 * (j + 8 * k2) + 2 * a + 1);
        const cpx p0 = cmul(f0, cpx{ wa.x, wa.y }), p1 = cmul(f1, cpx{ wb2.x, wb2.y });
        v[k2] = p0 + p1;
    }
    float2* __restrict__ pl = W.pil + g * 8 + a * 4;
    if (j == 7) pl[0] = make_float2(v[0].x, v[0].y);
    if (j == 5) pl[1] = make_float2(v[2].x, v[2].y);
    if (j == 3) pl[2] = make_float2(v[5].x, v[5].y);
    if (j == 1) pl[3] = make_float2(v[7].x, v[7].y);
    __syncwarp();
    cpx ps;
    {
        // HT: stream 0 uses PILOT_HT_2_1 {1,1,-1,-1}, stream 1 PILOT_HT_2_2 {1,-1,-1,1}; VHT: {1,1,1,-1} for both;
        // all rotate left once per symbol; polarity index starts at 3 (HT) / 4 (VHT) (lib/demod2_impl.cc:158,204)
        const bool vht = fmt == C8B_F_VHT;
        const float P = __ldg(&lut->pilotP[((vht ? 4 : 3) + sidx) % 127]);
        const int sh = sidx & 3;
        const uint32_t base0 = vht ? 0x8u : 0xCu, base1 = vht ? 0x8u : 0x6u;     // bit m set = pilot m is -1
        float re = 0.f, im = 0.f;
        const int binSlot[4] = { 2, 3, 0, 1 };                                   // bins 7,21,43,57 pair with d_pilot[2],[3],[0],[1]
#pragma unroll
        for (int st = 0; st < 2; st++) {
            const uint32_t base = st ? base1 : base0;
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const int m = binSlot[q];
                const float sgn = ((base >> ((m + sh) & 3)) & 1u) ? -P : P;
                const float2 sv = W.pil[g * 8 + st * 4 + q];
                const float2 rf = __ldg(w2 + 256 + st * 4 + m);
                const float tr = sv.x * sgn, ti = sv.y * sgn;                      // sig * pilot * P  (exact)
                const float pr = tr * rf.x - ti * rf.y, pi = tr * rf.y + ti * rf.x;
                re = (st == 0 && q == 0) ? pr : __fadd_rn(re, pr);
                im = (st == 0 && q == 0) ? pi : __fadd_rn(im, pi);
            }
        }
        const float inv = 1.0f / sqrtf(re * re + im * im);
        ps = { re * inv, -im * inv };
    }
    // soft bits of this half's stream through its deinterleaver and the stream parser, both in closed form (lut.h demapTab2)
    float* __restrict__ L = W.llr + g * ncbps;
    if (live) {
        const uint4 te = __ldg(reinterpret_cast<const uint4*>(lut->demapTab2[mi][a] + 8 * j));
        if (nbpsc == 1) demap2_symbol<1>(L, te, v, ps);
        else if (nbpsc == 2) demap2_symbol<2>(L, te, v, ps);
        else if (nbpsc == 4) demap2_symbol<4>(L, te, v, ps);
        else if (nbpsc == 6) demap2_symbol<6>(L, te, v, ps);
        else demap2_symbol<8>(L, te, v, ps);
    }
    bulk_store_fence();
    __syncwarp();
    const int wsym0 = sym0 + warp * SPW2;
    int nlive = nsym - wsym0;
    nlive = nlive < 0 ? 0 : (nlive > SPW2 ? SPW2 : nlive);
    const int nfl = nlive * ncbps;
    float* __restrict__ out = llrArena + llr_off + (int64_t)wsym0 * ncbps;
    if ((reinterpret_cast<uintptr_t>(out) & 15) == 0) {
        if (lane == 0 && nlive > 0) { bulk_store(out, W.llr, (uint32_t)nfl * 4u); bulk_store_drain(); }
    } else {
        for (int i = lane; i < nfl; i += 32) out[i] = W.llr[i];
    }
}

}  // namespace

void c8b_launch_demod2(const c8b_lut* lut, const float2* iq0, const float2* iq1, const int64_t* d_off, int nitems, int maxf, int maxSym,
                       const c8b_frame* frames, const float2* w2, float* llr, cudaStream_t st)
{
    if (nitems <= 0 || maxSym <= 0) return;
    const int per = (65535 / maxf) > 0 ? (65535 / maxf) : 1;   // items per launch (grid.y <= 65535 slots)
    for (int base = 0; base < nitems; base += per) {
        const int cnt = nitems - base < per ? nitems - base : per;
        dim3 grid((maxSym + SPB2 - 1) / SPB2, cnt * maxf);
        k_demod2<<<grid, DW * 32, 0, st>>>(lut, iq0, iq1, d_off + base, maxf, frames + (size_t)base * maxf, w2 + (size_t)base * maxf * 264, llr);
    }
}

void c8b_launch_demod(const c8b_lut* lut, const float2* iq, const int64_t* d_off, int nitems, int maxf, int maxSym, const c8b_frame* frames,
                      const float2* hinv, float* llr, cudaStream_t st)
{
    if (nitems <= 0 || maxSym <= 0) return;
    const int per = (65535 / maxf) > 0 ? (65535 / maxf) : 1;
    for (int base = 0; base < nitems; base += per) {
        const int cnt = nitems - base < per ? nitems - base : per;
        dim3 grid((maxSym + SPB - 1) / SPB, cnt * maxf);
        k_demod<<<grid, DW * 32, 0, st>>>(lut, iq, d_off + base, maxf, frames + (size_t)base * maxf, hinv + (size_t)base * maxf * 64, llr);
    }
}
