// k_viterbi.cu -- decode block on the GPU: depuncture + soft Viterbi (K=7, 64 states) + full
// traceback + descramble + A-MPDU walk + CRC-32.  One warp decodes TWO frames at a time (two
// independent add-compare-select chains interleaved for latency hiding), persistent CTAs.
//
// Replaces lib/decode_impl.cc:164-203 (vstb_init), :205-281 (vstb_update), :282-302 (vstb_end),
// :304-323 (descramble), :325-520 (packetAssemble).  Results are bit-exact with the reference for
// finite LLRs: path metrics are float32 sums formed in the reference's order (pre + tab[out],
// tab = {0, t1, t0, t1+t0}), ties go to the even predecessor (strict '>' with 2k visited first),
// the traceback starts in state 0 and covers the whole packet.
//
// Forward pass.  A trellis step maps old states (2k, 2k+1) to new states (k, k+32): one butterfly.
// Each lane keeps one butterfly's two old metrics in registers (x0 = metric[2k], x1 = metric[2k+1]),
// so the add-compare-select needs no communication; the two results (k, k+32) then have to become
// the (2k', 2k'+1) pair of some lane for the next step.  That is a swap of ONE lane-index bit with
// the register index -- one __shfl_xor per step -- and the lane bit that is swapped rotates with
// period 5 (phase P = t mod 5: the butterfly k of step t lives in lane rotl5(k, P)).
// Branch metrics: the step's table {0, t1, t0, t1+t0} is staged in shared memory (depunctured on
// load, 150 steps ahead) and each lane reads its two entries A = tab[c], B = tab[3-c] (c = encoder
// output of 2k --0--> k; the other three branches of the butterfly follow from both generator
// polynomials having taps at the newest and the oldest bit).
// Decisions: the sign of (even - odd) is shifted into a per-lane history word (one per result), so
// after 30 steps every lane owns 2x30 decisions: 256 B per warp and frame, one coalesced store.
// 16 SASS instructions per step and frame: 2 LDS, 6 FADD, 2 FMNMX, 2 SHF, 3 LOP3, 1 SHFL; the step is
// a ~100-cycle dependent chain (the shuffle alone ~40), hence two frames per warp.
//
// Traceback.  Runs in the same rotated domain: sig4 = 4*(2*lane + word) addresses the history word
// that holds the decision of the current state; a step is a shared-memory read, a rotate and three
// LOP3.  The packet is cut into 32 runs, one per lane, each started 4 groups (120 steps) early from
// state 0; run boundaries are verified against the neighbouring lane, a mismatch (paths not yet
// merged) falls back to the serial walk, so the result is always the reference's path.
// Decoded bits are packed LSB-first into 32-bit words, which makes the word array the PSDU bytes.
#include "common.cuh"

namespace {

constexpr int CH = C8B_VIT_CH;                 // 150 trellis steps per chunk
constexpr int GS = 30;                         // steps per decision group (one 30-bit history word per lane and result)
constexpr int NG = CH / GS;                    // 5 groups per chunk
constexpr int NW = C8B_VIT_WARPS;
constexpr int NLD = (CH + 31) / 32;            // table rows each lane fills per chunk
constexpr int WORDS = (C8B_DECODE_T_MAX + 31) / 32 + 8;  // decoded-bit words per warp (+ slack)
static_assert(CH % 5 == 0 && GS % 5 == 0 && CH % GS == 0, "phases must align with groups");
static_assert(((C8B_DECODE_T_MAX + CH - 1) / CH) * NG * 32 <= C8B_VIT_TPAD, "survivor scratch too small");

#ifndef C8B_VIT_UNROLL
#define C8B_VIT_UNROLL 6
#endif
constexpr int VUNROLL = C8B_VIT_UNROLL;          // 5-step iterations unrolled in the forward loop (6 = a whole 30-step group)
constexpr int TBW = 4;                         // traceback warm-up, in 30-step groups, before a lane's own segment
constexpr int U_BYTES = 9600;                  // union area: forward tables | traceback staging | decoded words
constexpr int ROW = 68;                        // words per lane row in the traceback staging (64 + pad: conflict-free 16-byte stores)
static_assert(2 * CH * 32 <= U_BYTES && 32 * ROW * 4 <= U_BYTES && WORDS * 4 <= U_BYTES, "union area too small");

// Per-warp shared memory (dynamic).  The union area is used, in turn, as
//   forward pass : float2 tab[2][CH][4] per step and class c the pair (tab[c], tab[3-c]) of {0, t1, t0, t1+t0}
//                  (lib/decode_impl.cc:231-234), one table per frame
//   traceback    : uint32 stage[32][ROW] decision words of the group each lane is walking, [lane][2*rho+h]
//   afterwards   : uint32 words[WORDS]  decoded bits packed LSB-first, descrambled in place -> PSDU bytes
struct __align__(16) WarpSmem {
    uint4 u[U_BYTES / 16];
    uint32_t scr[8];            // scrambler sequence, 160 bits
};

__device__ __forceinline__ int rotr5(int x, int p) { return ((x >> p) | (x << (5 - p))) & 31; }

// soft-bit indices of trellis step t for code rate cr; -1 = punctured position (metric 0.0f).
// Puncture patterns of lib/cloud80211phy.cc:1857-1860 in closed form.
__device__ __forceinline__ void depunc(int cr, int t, int& i0, int& i1)
{
    if (cr == C8B_CR_12) { i0 = 2 * t; i1 = 2 * t + 1; }
    else if (cr == C8B_CR_23) { int q = t >> 1, b = 3 * q; if (t & 1) { i0 = b + 2; i1 = -1; } else { i0 = b; i1 = b + 1; } }
    else if (cr == C8B_CR_34) {
        int q = t / 3, r = t - 3 * q, b = 4 * q;
        if (r == 0) { i0 = b; i1 = b + 1; } else if (r == 1) { i0 = b + 2; i1 = -1; } else { i0 = -1; i1 = b + 3; }
    } else {
        int q = t / 5, r = t - 5 * q, b = 6 * q;
        if (r == 0) { i0 = b; i1 = b + 1; }
        else if (r == 1) { i0 = b + 2; i1 = -1; }
        else if (r == 2) { i0 = -1; i1 = b + 3; }
        else if (r == 3) { i0 = b + 4; i1 = -1; }
        else { i0 = -1; i1 = b + 5; }
    }
}

// soft bits consumed by the first T steps, and per chunk of CH steps (CH is a multiple of every puncture period)
__device__ __forceinline__ int used_by(int cr, int T)
{
    if (cr == C8B_CR_12) return 2 * T;
    if (cr == C8B_CR_23) return 3 * (T >> 1) + ((T & 1) ? 2 : 0);
    if (cr == C8B_CR_34) { const int q = T / 3, r = T - 3 * q; return 4 * q + (r == 0 ? 0 : r + 1); }
    const int q = T / 5, r = T - 5 * q;
    return 6 * q + (r == 0 ? 0 : r + 1);
}
// packed chunk-relative indices of step s (0..CH-1): lo 16 bits -> t0, hi 16 bits -> t1, 0xffff = punctured
__device__ __forceinline__ uint32_t rel_pack(int cr, int s)
{
    int i0, i1;
    depunc(cr, s, i0, i1);
    return (uint32_t)(i0 & 0xffff) | ((uint32_t)(i1 & 0xffff) << 16);
}
// (t0,t1) of one step: base = first soft bit of the chunk, lim = soft bits the packet consumes (pad steps read 0)
__device__ __forceinline__ float2 load_pair(const float* __restrict__ llr, int base, int lim, uint32_t e)
{
    const int r0 = (int)(e & 0xffffu), r1 = (int)(e >> 16);
    const int i0 = base + r0, i1 = base + r1;
    float2 v;
    v.x = (r0 != 0xffff && i0 < lim) ? __ldg(llr + i0) : 0.0f;
    v.y = (r1 != 0xffff && i1 < lim) ? __ldg(llr + i1) : 0.0f;
    return v;
}
// per step, per butterfly class c: the pair (A, B) = (tab[c], tab[3-c]) of tab = {0, t1, t0, t1+t0}, so a lane gets both
// branch metrics with one 8-byte shared-memory read: {0,T3}, {t1,t0}, {t0,t1}, {T3,0}
__device__ __forceinline__ void put_tab(float4* __restrict__ row, float2 p)
{
    const float t3 = __fadd_rn(p.y, p.x);
    row[0] = make_float4(0.0f, t3, p.y, p.x);
    row[1] = make_float4(p.x, p.y, t3, 0.0f);
}

// one add-compare-select step at layout phase P.  (x0,x1) in: metrics of old states (2k,2k+1);
// out: metrics of (2k',2k'+1) for the next phase.  hLo/hHi: decision history of this lane.
template <int P>
__device__ __forceinline__ void acs_step(float& x0, float& x1, const float A, const float B, const uint32_t amask,
                                         uint32_t& hLo, uint32_t& hHi)
{
    const float eLo = __fadd_rn(x0, A), oLo = __fadd_rn(x1, B);   // into state k      (input 0)
    const float eHi = __fadd_rn(x0, B), oHi = __fadd_rn(x1, A);   // into state k + 32 (input 1)
    // odd predecessor wins only if strictly larger  <=>  (even - odd) is negative
    const uint32_t sLo = __float_as_uint(__fsub_rn(eLo, oLo)), sHi = __float_as_uint(__fsub_rn(eHi, oHi));
    const uint32_t y0 = __float_as_uint(fmaxf(eLo, oLo)), y1 = __float_as_uint(fmaxf(eHi, oHi));
    hLo = __funnelshift_l(sLo, hLo, 1);                            // (hLo << 1) | sign
    hHi = __funnelshift_l(sHi, hHi, 1);
    // swap lane bit P with the register index: lanes with the bit set send y0 and keep y1
    const uint32_t send = (y0 & amask) | (y1 & ~amask);
    const uint32_t recv = __shfl_xor_sync(0xffffffffu, send, 1 << P);
    x0 = __uint_as_float((recv & amask) | (y0 & ~amask));
    x1 = __uint_as_float((y1 & amask) | (recv & ~amask));
}

// traceback step for group-relative step index i (0..29); POS = ((i % 5) + 4) % 5.
// sig4 = 4*(2*rho + h): byte offset of the decision word inside the group's 256-byte block;
// the decision of step i sits at bit (GS-1-i) of that word.
__device__ __forceinline__ void tb_step(uint32_t& sig4, uint32_t& acc, const uint32_t* __restrict__ grp, const int i)
{
    const int POS = ((i % 5) + 4) % 5;
    const uint32_t M = 8u << POS;                                    // bit of rho[POS] inside sig4
    const uint32_t w = *reinterpret_cast<const uint32_t*>(reinterpret_cast<const char*>(grp) + sig4);
    acc = (acc << 1) | ((sig4 >> 2) & 1u);                           // decoded bit = h (input bit of the state entered)
    const uint32_t hn = (sig4 >> (POS + 1)) & 4u;                    // next h = old rho[POS], placed at bit 2
    const uint32_t pre = (sig4 & ~(M | 4u)) | hn;
    const uint32_t rot = __funnelshift_r(w, w, (GS - 1 - i - POS - 3) & 31);   // decision bit -> bit POS+3
    sig4 = pre | (rot & M);
}

// CRC-32 (boost::crc_32_type, lib/decode_impl.h:84): reflected 0x04C11DB7, init/xorout ~0.
// Lane-parallel: the message is cut into 64-byte segments aligned to its END; lane k runs the byte-wise
// table CRC over segments k and k+32 (register preset to ~0 only for the head segment), advances the
// result over the bytes that follow with the precomputed linear maps Z^(64*2^p), and the warp XORs the
// pieces.  Same value as the serial loop for every length (tests/test_gpu_decode.py).
__device__ __forceinline__ uint32_t crc_seg(const uint32_t* __restrict__ tab, const uint8_t* __restrict__ p, int n, uint32_t c)
{
    for (int i = 0; i < n; i++) c = tab[(c ^ p[i]) & 0xff] ^ (c >> 8);
    return c;
}
__device__ __forceinline__ uint32_t crc_advance(const uint32_t* __restrict__ zm, uint32_t c, int m)   // c * Z^(64 m)
{
#pragma unroll 1
    for (int p = 0; p < 6; p++) {
        if (!((m >> p) & 1)) continue;
        const uint32_t* __restrict__ col = zm + p * 32;
        uint32_t r = 0;
#pragma unroll 8
        for (int i = 0; i < 32; i++) r ^= ((c >> i) & 1u) ? col[i] : 0u;
        c = r;
    }
    return c;
}
__device__ __forceinline__ uint32_t crc32_warp(const uint32_t* __restrict__ tab, const uint32_t* __restrict__ zm,
                                               const uint8_t* __restrict__ p, int n, int lane)
{
    if (n <= 0) return 0u;                                       // ~(~0) : CRC of the empty message
    const int K = (n + 63) >> 6;                                 // segments, K <= 64 for n <= 4095 (A-MPDU: n < 4100)
    const int head = n - 64 * (K - 1);                           // bytes in segment 0 (1..64)
    uint32_t acc = 0;
    for (int k = lane; k < K; k += 32) {
        const int start = k == 0 ? 0 : head + 64 * (k - 1);
        const int len = k == 0 ? head : 64;
        uint32_t c = crc_seg(tab, p + start, len, k == 0 ? 0xffffffffu : 0u);
        acc ^= crc_advance(zm, c, K - 1 - k);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc ^= __shfl_xor_sync(0xffffffffu, acc, o);
    return ~acc;
}

// append one record [fmt][len lo][len hi][MPDU][mcs] to the frame's PDU area (all lanes cooperate)
__device__ __forceinline__ void emit_record(uint8_t* __restrict__ out, int& w, int cap, int& npdu, int fmt, int lenField,
                                            const uint8_t* __restrict__ body, int nbody, int mcs, int lane)
{
    const int rec = nbody + 4;
    if (w + rec > cap) return;
    uint8_t* o = out + w;
    if (lane == 0) { o[0] = (uint8_t)fmt; o[1] = (uint8_t)(lenField & 255); o[2] = (uint8_t)(lenField >> 8); o[3 + nbody] = (uint8_t)mcs; }
    for (int i = lane; i < nbody; i += 32) o[3 + i] = body[i];
    w += rec;
    npdu++;
}

struct Job {          // one frame's decode parameters (uniform across the warp)
    int f, T, cr, total, fmt, len, mcs, ampdu;
    const float* llr;
};

// reads frame f; returns T (0 = nothing to decode).  Writes the reject status of lib/decode_impl.cc:93-97.
__device__ __forceinline__ Job load_job(c8b_frame* __restrict__ frames, int f, int nframes, const float* __restrict__ llrArena,
                                        int64_t nllr, int64_t pduStride, bool lane0)
{
    Job j;
    j.f = f; j.T = 0; j.cr = 0; j.total = 0; j.fmt = 0; j.len = 0; j.mcs = 0; j.ampdu = 0; j.llr = llrArena;
    if (f >= nframes) return j;
    c8b_frame* fr = frames + f;
    const int status = fr->status;
    const int T = fr->trellis, total = fr->total, len = fr->len;
    const int64_t loff = fr->llr_off;
    j.cr = fr->cr & 3; j.total = total; j.fmt = fr->format; j.len = len; j.mcs = fr->mcs; j.ampdu = fr->ampdu;
    __syncwarp();
    if (lane0) { fr->npdu = 0; fr->pdu_bytes = 0; fr->pdu_off = (int64_t)f * pduStride; }
    if (status != C8B_ST_OK) return j;
    if (len > C8B_DECODE_B_MAX || T > C8B_DECODE_T_MAX) {          // lib/decode_impl.cc:93-97
        if (lane0) fr->status = C8B_ST_DECODE_RANGE;
        return j;
    }
    if (T <= 0 || total < 0 || loff < 0 || loff + total > nllr) return j;
    j.T = T; j.llr = llrArena + loff;
    return j;
}

#ifndef C8B_VIT_BLOCKS
#define C8B_VIT_BLOCKS 5
#endif
__global__ void __launch_bounds__(NW * 32, C8B_VIT_BLOCKS)
k_viterbi(const c8b_lut* __restrict__ lut, c8b_frame* __restrict__ frames, int nframes, const float* __restrict__ llrArena,
          int64_t nllr, uint2* __restrict__ survScratch, uint8_t* __restrict__ pdu, int64_t pduStride,
          uint8_t* __restrict__ scram, int64_t scramStride, unsigned* __restrict__ counter)
{
    extern __shared__ __align__(16) uint8_t dynsm[];
    WarpSmem* sm = reinterpret_cast<WarpSmem*>(dynsm);
    uint32_t* crcTab = reinterpret_cast<uint32_t*>(dynsm + NW * sizeof(WarpSmem));
    uint32_t* crcZ = crcTab + 256;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 256; i += NW * 32) crcTab[i] = lut->crc32tab[i];
    for (int i = threadIdx.x; i < 192; i += NW * 32) crcZ[i] = lut->crcZ[i / 32][i % 32];
    __syncthreads();
    WarpSmem& S = sm[warp];
    float4* __restrict__ Stab = reinterpret_cast<float4*>(S.u);        // [2 frames][CH]
    uint32_t* __restrict__ Sstage = reinterpret_cast<uint32_t*>(S.u);  // [32][64]
    uint32_t* __restrict__ Swords = reinterpret_cast<uint32_t*>(S.u);  // [WORDS]
    uint2* __restrict__ survW = survScratch + (size_t)(blockIdx.x * NW + warp) * C8B_VIT_WARP_SLOTS;
    const bool lane0 = lane == 0;

    // per-phase: index of this lane's A entry in the step table, and its side in the exchange
    int cA[5];
    uint32_t amask[5];
#pragma unroll
    for (int p = 0; p < 5; p++) {
        cA[p] = lut->bmClass[rotr5(lane, p)];
        uint32_t m = ((lane >> p) & 1) ? 0xffffffffu : 0u;
        asm volatile("" : "+r"(m));                                  // keep it a register mask (LOP3), not a predicate
        amask[p] = m;
    }

    for (;;) {
        int f0 = 0;
        if (lane0) f0 = (int)atomicAdd(counter, 2u);
        f0 = __shfl_sync(0xffffffffu, f0, 0);
        if (f0 >= nframes) break;
        Job jobs[2];
        jobs[0] = load_job(frames, f0, nframes, llrArena, nllr, pduStride, lane0);
        jobs[1] = load_job(frames, f0 + 1, nframes, llrArena, nllr, pduStride, lane0);
        const int TA = jobs[0].T, TB = jobs[1].T;
        const int nchA = (TA + CH - 1) / CH, nchB = (TB + CH - 1) / CH;
        const int nch = max(nchA, nchB);
        if (nch == 0) continue;

        // ---------------- forward pass, frames A and B interleaved ----------------
        {
            float xa0 = lane0 ? 0.0f : -1000000000000000.0f, xa1 = -1000000000000000.0f;   // lib/decode_impl.cc:171-176
            float xb0 = xa0, xb1 = xa1;
            float2 pfa[NLD], pfb[NLD];
            // chunk-relative soft-bit indices of this lane's table rows (same for every chunk)
            uint32_t rela[NLD], relb[NLD];
            const int nrawA = used_by(jobs[0].cr, CH), nrawB = used_by(jobs[1].cr, CH);
            const int limA = min(jobs[0].total, used_by(jobs[0].cr, TA)), limB = min(jobs[1].total, used_by(jobs[1].cr, TB));
#pragma unroll
            for (int j = 0; j < NLD; j++) {
                const int sidx = min(lane + 32 * j, CH - 1);
                rela[j] = rel_pack(jobs[0].cr, sidx);
                relb[j] = rel_pack(jobs[1].cr, sidx);
            }
            __syncwarp();
#pragma unroll
            for (int j = 0; j < NLD; j++)
                if (lane + 32 * j < CH) {
                    put_tab(Stab + 2 * (lane + 32 * j), load_pair(jobs[0].llr, 0, limA, rela[j]));
                    put_tab(Stab + 2 * (CH + lane + 32 * j), load_pair(jobs[1].llr, 0, limB, relb[j]));
                }
            __syncwarp();
            const float2* __restrict__ tb = reinterpret_cast<const float2*>(Stab);   // [frame][step][class] -> (A, B)
            for (int c = 0; c < nch; c++) {
                const bool more = c + 1 < nch;
                if (more) {
#pragma unroll
                    for (int j = 0; j < NLD; j++) {
                        pfa[j] = load_pair(jobs[0].llr, (c + 1) * nrawA, limA, rela[j]);
                        pfb[j] = load_pair(jobs[1].llr, (c + 1) * nrawB, limB, relb[j]);
                    }
                }
                const float2* p0 = tb + 0 + cA[0], *p1 = tb + 4 + cA[1], *p2 = tb + 8 + cA[2], *p3 = tb + 12 + cA[3], *p4 = tb + 16 + cA[4];
                uint2* __restrict__ sgA = survW + (size_t)c * (NG * 32) + lane;
                uint2* __restrict__ sgB = sgA + C8B_VIT_TPAD;
                const bool stA = c < nchA, stB = c < nchB;
#pragma unroll 1
                for (int g = 0; g < NG; g++) {
                    uint32_t haLo = 0, haHi = 0, hbLo = 0, hbHi = 0;
#pragma unroll VUNROLL
                    for (int i5 = 0; i5 < GS / 5; i5++) {           // 30 steps per group, 5 layout phases per iteration
                        constexpr int ob = CH * 4;                  // frame B's table sits CH steps (x 4 classes) further
                        float2 ab;
                        ab = p0[0];  acs_step<0>(xa0, xa1, ab.x, ab.y, amask[0], haLo, haHi);
                        ab = p0[ob]; acs_step<0>(xb0, xb1, ab.x, ab.y, amask[0], hbLo, hbHi);
                        ab = p1[0];  acs_step<1>(xa0, xa1, ab.x, ab.y, amask[1], haLo, haHi);
                        ab = p1[ob]; acs_step<1>(xb0, xb1, ab.x, ab.y, amask[1], hbLo, hbHi);
                        ab = p2[0];  acs_step<2>(xa0, xa1, ab.x, ab.y, amask[2], haLo, haHi);
                        ab = p2[ob]; acs_step<2>(xb0, xb1, ab.x, ab.y, amask[2], hbLo, hbHi);
                        ab = p3[0];  acs_step<3>(xa0, xa1, ab.x, ab.y, amask[3], haLo, haHi);
                        ab = p3[ob]; acs_step<3>(xb0, xb1, ab.x, ab.y, amask[3], hbLo, hbHi);
                        ab = p4[0];  acs_step<4>(xa0, xa1, ab.x, ab.y, amask[4], haLo, haHi);
                        ab = p4[ob]; acs_step<4>(xb0, xb1, ab.x, ab.y, amask[4], hbLo, hbHi);
                        p0 += 20; p1 += 20; p2 += 20; p3 += 20; p4 += 20;
                    }
                    if (stA) sgA[g * 32] = make_uint2(haLo, haHi);
                    if (stB) sgB[g * 32] = make_uint2(hbLo, hbHi);
                }
                __syncwarp();
                if (more) {
#pragma unroll
                    for (int j = 0; j < NLD; j++)
                        if (lane + 32 * j < CH) { put_tab(Stab + 2 * (lane + 32 * j), pfa[j]); put_tab(Stab + 2 * (CH + lane + 32 * j), pfb[j]); }
                }
                __syncwarp();
            }
        }

        // ---------------- per frame: traceback, descramble, assemble ----------------
#pragma unroll 1
        for (int which = 0; which < 2; which++) {
            const Job jb = jobs[which];
            const int T = jb.T;
            if (T <= 0) continue;
            const int f = jb.f, fmt = jb.fmt, len = jb.len, mcs = jb.mcs, ampdu = jb.ampdu;
            c8b_frame* fr = frames + f;
            const uint2* __restrict__ survG = survW + (size_t)which * C8B_VIT_TPAD;

            // traceback (lib/decode_impl.cc:282-302), final state 0.  The packet is cut into 32 runs of groups,
            // one per lane; lane l starts TBW groups above its run from an arbitrary state (0); by the time it
            // enters its own run its path has (almost surely) merged with the true one.  The run boundaries are
            // then CHECKED: lane l's state on entering its run must equal the state lane l+1 reached on leaving
            // its own (the top lane starts from the true end).  If every boundary agrees the 32 pieces are
            // exactly the reference's path; otherwise the packet is walked serially.
            // gbits[G] = 30 decoded bits of group G (bit i = step 30G+i); lives behind the warp's survivor scratch
            uint32_t* __restrict__ gbits = reinterpret_cast<uint32_t*>(survW + 2 * C8B_VIT_TPAD);
            const int NGR = (T + GS - 1) / GS;
            bool merged;
            {
                const int per = (NGR + 31) >> 5;                     // groups per lane
                const int nseg = (NGR + per - 1) / per;              // lanes that own a run
                const int gs = lane * per, ge = min(gs + per, NGR);
                const bool has = gs < NGR;
                const int gtop = has ? min(ge + TBW, NGR) - 1 : -1;  // first (highest) group this lane walks
                uint32_t sig4 = 0, sigEnd = 0, sigIn = 0;
                uint32_t* __restrict__ myrow = Sstage + lane * ROW;
                for (int r = 0; r < per + TBW; r++) {
                    const int G = gtop - r;
                    const bool act = has && G >= gs;
                    if (act) {                                       // copy the 256-byte row of group G next to the lane (lane-private)
                        const uint4* __restrict__ src = reinterpret_cast<const uint4*>(survG + (size_t)G * 32);
                        uint4* __restrict__ dst = reinterpret_cast<uint4*>(myrow);
#pragma unroll
                        for (int h8 = 0; h8 < 16; h8 += 8) {
                            uint4 v[8];
#pragma unroll
                            for (int k = 0; k < 8; k++) v[k] = src[h8 + k];
#pragma unroll
                            for (int k = 0; k < 8; k++) dst[h8 + k] = v[k];
                        }
                        if (G == ge - 1) sigEnd = sig4;              // state at the upper boundary of the run
                        const int lim = T - G * GS;                  // steps left in this group (>= GS unless it is the last)
                        uint32_t acc = 0;
                        if (lim >= GS) {
#pragma unroll
                            for (int i = GS - 1; i >= 0; i--) tb_step(sig4, acc, myrow, i);
                        } else {
                            for (int i = lim - 1; i >= 0; i--) tb_step(sig4, acc, myrow, i);
                        }
                        if (G < ge) gbits[G] = acc;
                        if (G == gs) sigIn = sig4;                   // state at the lower boundary of the run
                    }
                }
                const uint32_t above = __shfl_down_sync(0xffffffffu, sigIn, 1);
                const bool ok = !(has && lane < nseg - 1) || sigEnd == above;
                merged = __all_sync(0xffffffffu, ok);
            }
            if (!merged) {                                           // serial walk, all lanes redundantly
                uint32_t sig4 = 0;
                for (int G = NGR - 1; G >= 0; G--) {
                    __syncwarp();
                    *reinterpret_cast<uint2*>(&Sstage[lane * 2]) = survG[(size_t)G * 32 + lane];   // row 0 of the staging area
                    __syncwarp();
                    uint32_t acc = 0;
                    const int lim = min(GS, T - G * GS);
                    for (int i = lim - 1; i >= 0; i--) tb_step(sig4, acc, Sstage, i);
                    if (lane0) gbits[G] = acc;
                }
            }
            __syncwarp();
            // repack 30-bit groups into the 32-bit LSB-first word stream (word w = steps 32w .. 32w+31)
            const int nwords = (T + 31) >> 5;
            for (int w = lane; w < nwords; w += 32) {
                const int b0 = 32 * w, G = b0 / GS, off = b0 - G * GS;
                uint32_t v = gbits[G] >> off;                        // GS-off bits
                if (G + 1 < NGR) v |= gbits[G + 1] << (GS - off);
                if (2 * GS - off < 32 && G + 2 < NGR) v |= gbits[G + 2] << (2 * GS - off);
                Swords[w] = v;
            }
            __syncwarp();
            if (scram != nullptr) {
                uint8_t* so = scram + (size_t)f * scramStride;
                for (int i = lane; i < T && i < scramStride; i += 32) so[i] = (uint8_t)((Swords[i >> 5] >> (i & 31)) & 1u);
            }

            // descramble (lib/decode_impl.cc:304-323)
            {
                const uint32_t w0 = Swords[0];
                int st = 0;
#pragma unroll
                for (int i = 0; i < 7; i++) st |= (int)((w0 >> i) & 1u) << (6 - i);
                for (int wq = 0; wq < 5; wq++) {
                    uint32_t q = 0;
#pragma unroll 8
                    for (int b = 0; b < 32; b++) {
                        const int fb = ((st >> 6) ^ (st >> 3)) & 1;
                        st = ((st << 1) & 0x7e) | fb;
                        q |= (uint32_t)fb << b;
                    }
                    if (lane0) S.scr[wq] = q;
                }
                __syncwarp();
                for (int w = lane; w < nwords; w += 32) {
                    uint32_t v = Swords[w];
                    if (w == 0) v = (v ^ (S.scr[0] << 7)) & ~0x7fu;
                    else {
                        const int o = (32 * w - 7) % 127;
                        v ^= __funnelshift_r(S.scr[o >> 5], S.scr[(o >> 5) + 1], o & 31);
                    }
                    Swords[w] = v;
                }
                __syncwarp();
            }

            // packetAssemble (lib/decode_impl.cc:325-520)
            {
                const uint8_t* __restrict__ by = reinterpret_cast<const uint8_t*>(Swords);
                uint8_t* out = pdu + (size_t)f * pduStride;
                const int cap = (int)min(pduStride, (int64_t)0x7fffffff);
                int npdu = 0, w = 0;
                if (fmt == C8B_F_VHT) {
                    int procd = 16;
                    if (procd < T) {
                        int bp = 2;                                  // byte offset of the next delimiter
                        int tl = 0;                                  // NOT reset per subframe (:336)
                        while (true) {
                            procd += 32;
                            if (procd > T) break;
                            const int d0 = by[bp], d1 = by[bp + 1];
                            const int eof = d0 & 1;
                            tl |= ((d0 >> 2) & 1) << 12;
                            tl |= ((d0 >> 3) & 1) << 13;
                            tl |= (d0 >> 4) | (d1 << 4);
                            const int padded = (tl / 4 + ((tl % 4) != 0)) * 4;   // bytes
                            procd += padded * 8;
                            if (procd > T) break;
                            bp += 4;
                            const uint32_t crc = crc32_warp(crcTab, crcZ, by + bp, tl, lane);
                            if (crc == 558161692u) {
                                emit_record(out, w, cap, npdu, fmt, tl, by + bp, tl, mcs, lane);
                                tl += 4;                             // :415, carried into the next subframe
                            }
                            bp += padded;
                            if (eof) break;
                        }
                    }
                } else if (!ampdu) {
                    if (len >= 0 && 16 + 8 * len <= 32 * nwords) {
                        const uint32_t crc = crc32_warp(crcTab, crcZ, by + 2, len, lane);
                        if (crc == 558161692u) emit_record(out, w, cap, npdu, fmt, len, by + 2, len, mcs, lane);
                    }
                }
                if (lane0) { fr->npdu = npdu; fr->pdu_bytes = w; }
            }
            __syncwarp();
        }
    }
}

}  // namespace

int c8b_viterbi_max_grid(int num_sm) { return num_sm * C8B_VIT_BLOCKS; }

// per-device opt-in to > 48 KB of dynamic shared memory (see c8b_viterbi_tp_prepare)
cudaError_t c8b_viterbi_prepare(void)
{
    return cudaFuncSetAttribute(k_viterbi, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(NW * sizeof(WarpSmem) + 1024 + 768));
}

void c8b_launch_viterbi(const c8b_lut* d_lut, c8b_frame* d_frames, int nframes, const float* d_llr, int64_t nllr, uint2* d_surv,
                        int nwarps_alloc, uint8_t* d_pdu, int64_t pdu_stride, uint8_t* d_scram, int64_t scram_stride,
                        unsigned* d_counter, int grid, cudaStream_t st)
{
    if (nframes <= 0) return;
    int need = (nframes + 2 * NW - 1) / (2 * NW);
    if (grid > need) grid = need;
    if (grid * NW > nwarps_alloc) grid = nwarps_alloc / NW;
    const size_t smem = NW * sizeof(WarpSmem) + 1024 + 768;
    cudaMemsetAsync(d_counter, 0, sizeof(unsigned), st);
    k_viterbi<<<grid, NW * 32, smem, st>>>(d_lut, d_frames, nframes, d_llr, nllr, d_surv, d_pdu, pdu_stride, d_scram, scram_stride,
                                          d_counter);
}

// VHT NDP (lib/decode_impl.cc:100-121: v_trellis == 0): the decode block turns the tag "mu2x1chan" -- here the 128 complex
// samples the header kernel left at the frame's place in the LLR arena -- into the channel report
// [C8P_F_VHT_CHAN = 20][len lo][len hi][128 x (re, im) float32], len = 1024.  Runs after the Viterbi kernel (which clears
// npdu / pdu_bytes of every slot); one thread per frame slot, NDP frames are rare.
namespace {
__global__ void k_ndp(c8b_frame* __restrict__ frames, int nframes, const float* __restrict__ llr, int64_t nllr, uint8_t* __restrict__ pdu,
                      int64_t pduStride)
{
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= nframes) return;
    c8b_frame* fr = frames + f;
    if (fr->status != C8B_ST_NDP || pduStride < 1027) return;
    const int64_t lo = fr->llr_off;
    if (lo < 0 || lo + 256 > nllr) return;
    uint8_t* __restrict__ p = pdu + (size_t)f * pduStride;
    const uint8_t* __restrict__ s = reinterpret_cast<const uint8_t*>(llr + lo);
    p[0] = 20; p[1] = 1024 % 256; p[2] = 1024 / 256;
    for (int k = 0; k < 1024; k++) p[3 + k] = s[k];
    fr->npdu = 1; fr->pdu_bytes = 1027; fr->pdu_off = (int64_t)f * pduStride;
}
}  // namespace

void c8b_launch_ndp(c8b_frame* frames, int nframes, const float* llr, int64_t nllr, uint8_t* pdu, int64_t pduStride, cudaStream_t st)
{
    if (nframes <= 0) return;
    k_ndp<<<(nframes + 255) / 256, 256, 0, st>>>(frames, nframes, llr, nllr, pdu, pduStride);
}
