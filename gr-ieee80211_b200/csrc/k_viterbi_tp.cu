// k_viterbi_tp.cu -- decode block, throughput variant: ONE THREAD PER FRAME.
//
// Same function and same results as k_viterbi.cu (lib/decode_impl.cc:164-520: depuncture, 64-state soft
// Viterbi with float32 metrics in the reference's operation order, full traceback from state 0, descramble,
// A-MPDU walk, CRC-32), organised for large batches: a thread keeps all 64 path metrics of its frame in
// registers, so a trellis step is 32 butterflies of straight-line code with NO shuffles and NO shared-memory
// traffic in the dependent chain -- the branch-metric class of every butterfly is a compile-time constant
// (9 instructions per butterfly: <=4 FADD, 2 FADD for the decision signs, 2 FMNMX, 2 SHF), 32 independent
// chains per thread.  k_viterbi.cu (one warp per frame pair, 15 instr per lane and step behind a ~100-cycle
// shuffle chain) remains the low-latency path for small batches; ctx.cu picks by batch size.
//
// Data movement.  A warp owns 32 frames.  Soft bits are staged per 30-step chunk: for each of its frames the
// warp reads the chunk's soft bits coalesced, depunctures them (lib/cloud80211phy.cc:1857-1860 patterns in
// closed form) and writes (t0,t1) pairs into a shared-memory row per frame; lane j then walks row j.
// Survivors: one 64-bit decision word per thread and step, stored [step][thread] so a warp writes 256 B
// contiguous; the traceback reads them back the same way.  Decoded bits go to a [word][thread] scratch,
// are descrambled there and walked by the same thread for the A-MPDU / CRC-32 / PDU output.
#include "common.cuh"

namespace {

constexpr int TPB = 128;                       // threads (= frames) per CTA (224 x 2 CTAs = 14 warps per SM at 144 registers spills: slower)
#ifndef C8B_TP_SURV_WARP
#define C8B_TP_SURV_WARP 0
#endif
constexpr int SVS = C8B_TP_SURV_WARP ? 32 : TPB;   // survivor words between two trellis steps of one thread
#ifndef C8B_TP_UNROLL
#define C8B_TP_UNROLL 2                        // trellis steps per trip of the forward loop (divides CS, even)
#endif
constexpr int CS = 30;                         // trellis steps per staged chunk (multiple of every puncture period and of 2)
constexpr int ROWF2 = CS + 1;                  // float2 per shared-memory row (odd -> conflict-free 8-byte column walks)

// encoder output class of the transition 2k --0--> k (same definition as lut.cc bmClass), compile time
__host__ __device__ constexpr int parity7(int v) { return (v ^ (v >> 1) ^ (v >> 2) ^ (v >> 3) ^ (v >> 4) ^ (v >> 5) ^ (v >> 6)) & 1; }
__host__ __device__ constexpr int bm_class(int k)
{
    int s6 = 2 * k, reg = 0;
    for (int q = 0; q < 6; q++) reg |= ((s6 >> (5 - q)) & 1) << (q + 1);
    return parity7(reg & 0155) * 2 + parity7(reg & 0117);
}

// one trellis step: m (metrics of the 64 old states) -> n.  wlo/whi: decision bits of new states 0..31 / 32..63.
// tab = {0, t1, t0, t1+t0}; pre + tab[0] is pre itself (adding +0.0f is exact), so those adds are skipped.
template <int K>
__device__ __forceinline__ void butterfly(const float (&m)[64], float (&n)[64], const float t0, const float t1, const float t3,
                                          uint32_t& wlo, uint32_t& whi)
{
    constexpr int c = bm_class(K);
    const float e = m[2 * K], o = m[2 * K + 1];
    float eLo, oLo, eHi, oHi;                                       // A = tab[c] on (e->lo, o->hi), B = tab[3-c] on (o->lo, e->hi)
    if (c == 0) { eLo = e; oLo = __fadd_rn(o, t3); eHi = __fadd_rn(e, t3); oHi = o; }
    else if (c == 1) { eLo = __fadd_rn(e, t1); oLo = __fadd_rn(o, t0); eHi = __fadd_rn(e, t0); oHi = __fadd_rn(o, t1); }
    else if (c == 2) { eLo = __fadd_rn(e, t0); oLo = __fadd_rn(o, t1); eHi = __fadd_rn(e, t1); oHi = __fadd_rn(o, t0); }
    else { eLo = __fadd_rn(e, t3); oLo = o; eHi = e; oHi = __fadd_rn(o, t3); }
    // odd predecessor wins only if strictly larger  <=>  (even - odd) is negative (ties keep the even one)
    wlo = __funnelshift_l(__float_as_uint(__fsub_rn(eLo, oLo)), wlo, 1);
    whi = __funnelshift_l(__float_as_uint(__fsub_rn(eHi, oHi)), whi, 1);
    n[K] = fmaxf(eLo, oLo);
    n[K + 32] = fmaxf(eHi, oHi);
}

template <int K>
struct Bf {
    static __device__ __forceinline__ void run(const float (&m)[64], float (&n)[64], float t0, float t1, float t3, uint32_t& wlo, uint32_t& whi)
    {
        butterfly<K>(m, n, t0, t1, t3, wlo, whi);                   // K = 31 first: its bit ends up at position 31
        Bf<K - 1>::run(m, n, t0, t1, t3, wlo, whi);
    }
};
template <>
struct Bf<-1> {
    static __device__ __forceinline__ void run(const float (&)[64], float (&)[64], float, float, float, uint32_t&, uint32_t&) {}
};

__device__ __forceinline__ uint2 acs(const float (&m)[64], float (&n)[64], const float2 tt)
{
    uint32_t wlo = 0, whi = 0;
    Bf<31>::run(m, n, tt.x, tt.y, __fadd_rn(tt.y, tt.x), wlo, whi);
    return make_uint2(wlo, whi);
}

// chunk-relative soft-bit indices of step s for code rate cr; -1 = punctured (same closed forms as k_viterbi.cu)
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
__device__ __forceinline__ int used_by(int cr, int T)
{
    if (cr == C8B_CR_12) return 2 * T;
    if (cr == C8B_CR_23) return 3 * (T >> 1) + ((T & 1) ? 2 : 0);
    if (cr == C8B_CR_34) { const int q = T / 3, r = T - 3 * q; return 4 * q + (r == 0 ? 0 : r + 1); }
    const int q = T / 5, r = T - 5 * q;
    return 6 * q + (r == 0 ? 0 : r + 1);
}

// Scratch per resident CTA, in global memory:
//   surv  [C8B_DECODE_T_MAX + CS][TPB] uint2     decision words, one per trellis step and thread
//   words [(C8B_DECODE_T_MAX + 63) / 32][TPB]    decoded bits, LSB first; descrambled in place

__device__ __forceinline__ uint32_t get_byte(const uint32_t* __restrict__ words, int i)
{
    return (words[(size_t)(i >> 2) * TPB] >> (8 * (i & 3))) & 0xffu;
}

// CRC-32 (boost::crc_32_type, lib/decode_impl.h:84) over bytes [start, start+n) of the thread's word column
__device__ uint32_t crc32_words(const uint32_t* __restrict__ tab, const uint32_t* __restrict__ words, int start, int n)
{
    uint32_t c = 0xffffffffu;
    int i = start;
    const int end = start + n;
    for (; i < end && (i & 3); i++) c = tab[(c ^ get_byte(words, i)) & 0xff] ^ (c >> 8);
    for (; i + 4 <= end; i += 4) {
        const uint32_t w = words[(size_t)(i >> 2) * TPB];
        c = tab[(c ^ w) & 0xff] ^ (c >> 8);
        c = tab[(c ^ (w >> 8)) & 0xff] ^ (c >> 8);
        c = tab[(c ^ (w >> 16)) & 0xff] ^ (c >> 8);
        c = tab[(c ^ (w >> 24)) & 0xff] ^ (c >> 8);
    }
    for (; i < end; i++) c = tab[(c ^ get_byte(words, i)) & 0xff] ^ (c >> 8);
    return ~c;
}

// [fmt][len lo][len hi][MPDU][mcs] appended to the frame's PDU area by one thread
__device__ void emit_record_t(uint8_t* __restrict__ out, int& w, int cap, int& npdu, int fmt, int lenField, const uint32_t* __restrict__ words,
                              int start, int nbody, int mcs)
{
    const int rec = nbody + 4;
    if (w + rec > cap) return;
    uint8_t* o = out + w;
    o[0] = (uint8_t)fmt; o[1] = (uint8_t)(lenField & 255); o[2] = (uint8_t)(lenField >> 8);
    int i = 0;
    // bytes until the destination is word aligned, then 4 bytes per store (source words funnel-shifted into place)
    for (; i < nbody && (((uintptr_t)(o + 3 + i)) & 3); i++) o[3 + i] = (uint8_t)get_byte(words, start + i);
    for (; i + 4 <= nbody; i += 4) {
        const int k = start + i;
        const uint32_t w0 = words[(size_t)(k >> 2) * TPB], w1 = words[(size_t)((k >> 2) + 1) * TPB];
        *reinterpret_cast<uint32_t*>(o + 3 + i) = __funnelshift_r(w0, w1, 8 * (k & 3));
    }
    for (; i < nbody; i++) o[3 + i] = (uint8_t)get_byte(words, start + i);
    o[3 + nbody] = (uint8_t)mcs;
    w += rec;
    npdu++;
}

#ifndef C8B_TP_CTAS
#define C8B_TP_CTAS 3                          // CTAs per SM: 3 (168 registers per thread, no spills) or 4 (128 registers; needs C8B_TP_RING)
#endif
#ifndef C8B_TP_RING
#define C8B_TP_RING (C8B_TP_CTAS > 3)          // 1: soft bits staged through ONE CS-step ring per lane, a third (SEG steps) at a time,
#endif                                         //    instead of two CS-step buffers: half the shared memory, so that 4 CTAs fit an SM
constexpr int SEG = 10;                        // steps per ring segment (even; CS = 3 segments)
constexpr int NBUF = C8B_TP_RING ? 1 : 2;
constexpr int TBK = C8B_TP_CTAS > 3 ? 16 : 32; // decision words the traceback keeps in flight per block (registers)
__global__ void __launch_bounds__(TPB, C8B_TP_CTAS)
k_viterbi_tp(const c8b_lut* __restrict__ lut, c8b_frame* __restrict__ frames, int nframes, const float* __restrict__ llrArena,
             int64_t nllr, uint2* __restrict__ survAll, uint32_t* __restrict__ wordsAll, size_t survPerCta, size_t wordsPerCta,
             uint8_t* __restrict__ pdu, int64_t pduStride, uint8_t* __restrict__ scram, int64_t scramStride)
{
    extern __shared__ __align__(16) uint8_t dynsm[];
    float2 (*pairs)[NBUF][32 * ROWF2] = reinterpret_cast<float2 (*)[NBUF][32 * ROWF2]>(dynsm);   // [warp][buffer][frame row][step]
    uint32_t* crcTab = reinterpret_cast<uint32_t*>(dynsm + sizeof(float2) * (TPB / 32) * NBUF * 32 * ROWF2);
    uint32_t* relTab = crcTab + 256;                                 // [4 code rates][32]
    for (int i = threadIdx.x; i < 256; i += TPB) crcTab[i] = lut->crc32tab[i];
    for (int i = threadIdx.x; i < 128; i += TPB) {
        int a0, a1;
        depunc(i >> 5, (i & 31) < CS ? (i & 31) : 0, a0, a1);
        relTab[i] = (uint32_t)(a0 & 0xffff) | ((uint32_t)(a1 & 0xffff) << 16);
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#if C8B_TP_SURV_WARP
    // every warp owns a contiguous survivor stream [step][lane]: its traceback reads 256-byte rows back to back
    uint2* __restrict__ surv = survAll + (size_t)blockIdx.x * survPerCta + (size_t)(threadIdx.x >> 5) * (survPerCta / (TPB / 32)) + (threadIdx.x & 31);
#else
    uint2* __restrict__ surv = survAll + (size_t)blockIdx.x * survPerCta + threadIdx.x;        // [t * TPB]
#endif
    uint32_t* __restrict__ words = wordsAll + (size_t)blockIdx.x * wordsPerCta + threadIdx.x;  // [w * TPB]

    for (int g = blockIdx.x; g * TPB < nframes; g += gridDim.x) {
        const int f = g * TPB + threadIdx.x;
        // ---- this thread's frame ----
        int T = 0, cr = 0, total = 0, fmt = 0, len = 0, mcs = 0, ampdu = 0;
        const float* llr = llrArena;
        if (f < nframes) {
            c8b_frame* fr = frames + f;
            const int status = fr->status;
            const int t = fr->trellis;
            const int64_t loff = fr->llr_off;
            cr = fr->cr & 3; total = fr->total; fmt = fr->format; len = fr->len; mcs = fr->mcs; ampdu = fr->ampdu;
            fr->npdu = 0; fr->pdu_bytes = 0; fr->pdu_off = (int64_t)f * pduStride;
            if (status == C8B_ST_OK) {
                if (len > C8B_DECODE_B_MAX || t > C8B_DECODE_T_MAX) fr->status = C8B_ST_DECODE_RANGE;     // lib/decode_impl.cc:93-97
                else if (t > 0 && total >= 0 && loff >= 0 && loff + total <= nllr) { T = t; llr = llrArena + loff; }
            }
        }
        const int lim = min(total, used_by(cr, T));                  // soft bits this packet consumes (pad steps read 0)
        const int nraw = used_by(cr, CS);                            // soft bits per chunk of this frame
        int Tmax = T;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) Tmax = max(Tmax, __shfl_xor_sync(0xffffffffu, Tmax, o));
        if (Tmax == 0) continue;                                     // warp-uniform
        const int nch = (Tmax + CS - 1) / CS;

        // stage chunk c: every lane copies the (t0,t1) pairs of ITS frame into its shared-memory row with cp.async
        // (LDGSTS): the copies of the next chunk run under the butterflies of the current one and hold no registers.
        // relTab[cr][s] = chunk-relative soft-bit indices of step s (i0 | i1 << 16, 0xffff = punctured).  A chunk is a whole
        // number of puncture periods, so the punctured slots of a row are the same in every chunk: they are zeroed once per
        // frame and never copied; a position past the end of the packet is a zero-fill copy (src-size 0).  The table entries
        // of ten steps are fetched together (LDS.64) ahead of their twenty copies -- one dependent LDS per step stalled the
        // warp for a fifth of the kernel.
        {
#pragma unroll
            for (int b = 0; b < NBUF; b++) {
                float2* r0 = pairs[warp][b] + lane * ROWF2;
#pragma unroll
                for (int i = 0; i < CS; i++) r0[i] = make_float2(0.f, 0.f);
            }
            __syncwarp();
        }
        auto stage = [&](int c, int buf) {
            const uint32_t d0 = (uint32_t)__cvta_generic_to_shared(pairs[warp][buf] + lane * ROWF2);
            const float* __restrict__ lb = llr + (size_t)c * nraw;
            const int rem = lim - c * nraw;                          // soft bits left from the start of this chunk
            const uint2* __restrict__ rt = reinterpret_cast<const uint2*>(relTab + cr * 32);
            if (rem >= nraw) {                                       // the whole chunk lies inside the packet: no end test per copy
#pragma unroll
                for (int b = 0; b < CS; b += 10) {
                    uint2 e[5];
#pragma unroll
                    for (int k = 0; k < 5; k++) e[k] = rt[b / 2 + k];
#pragma unroll
                    for (int k = 0; k < 10; k++) {
                        const uint32_t ev = (k & 1) ? e[k >> 1].y : e[k >> 1].x;
                        const uint32_t i0 = ev & 0xffffu, i1 = ev >> 16;
                        if (i0 != 0xffffu) asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d0 + 8 * (b + k)), "l"(lb + i0));
                        if (i1 != 0xffffu) asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d0 + 8 * (b + k) + 4), "l"(lb + i1));
                    }
                }
                asm volatile("cp.async.commit_group;" ::: "memory");
                return;
            }
#pragma unroll
            for (int b = 0; b < CS; b += 10) {
                uint2 e[5];
#pragma unroll
                for (int k = 0; k < 5; k++) e[k] = rt[b / 2 + k];
#pragma unroll
                for (int k = 0; k < 10; k++) {
                    const uint32_t ev = (k & 1) ? e[k >> 1].y : e[k >> 1].x;
                    const int i0 = (int)(ev & 0xffffu), i1 = (int)(ev >> 16);
                    const bool in0 = i0 < rem, in1 = i1 < rem;
                    if (i0 != 0xffff)
                        asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(d0 + 8 * (b + k)), "l"(lb + (in0 ? i0 : 0)), "r"(in0 ? 4 : 0));
                    if (i1 != 0xffff)
                        asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(d0 + 8 * (b + k) + 4), "l"(lb + (in1 ? i1 : 0)), "r"(in1 ? 4 : 0));
                }
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        };
        auto stage_wait = [&]() { asm volatile("cp.async.wait_group 0;" ::: "memory"); __syncwarp(); };

        // ---------------- forward pass ----------------
        float m[64], n[64];
#pragma unroll
        for (int i = 0; i < 64; i++) m[i] = -1000000000000000.0f;     // lib/decode_impl.cc:171-176
        m[0] = 0.0f;
#if C8B_TP_RING
        // ring: segment q = steps [q SEG, q SEG + SEG) lives in third q % 3 of the lane's row; its copies are issued while
        // segment q - 1 is computed (the third they overwrite was consumed by segment q - 2)
        auto stage_seg = [&](int q) {
            const int j = q % 3, c = q / 3;
            const uint32_t d0 = (uint32_t)__cvta_generic_to_shared(pairs[warp][0] + lane * ROWF2) + 8 * SEG * j;
            const float* __restrict__ lb = llr + (size_t)c * nraw;
            const int rem = lim - c * nraw;                          // soft bits left from the start of this CS-step period
            const uint2* __restrict__ rt = reinterpret_cast<const uint2*>(relTab + cr * 32) + (SEG / 2) * j;
            uint2 e[SEG / 2];
#pragma unroll
            for (int k = 0; k < SEG / 2; k++) e[k] = rt[k];
#pragma unroll
            for (int k = 0; k < SEG; k++) {
                const uint32_t ev = (k & 1) ? e[k >> 1].y : e[k >> 1].x;
                const int i0 = (int)(ev & 0xffffu), i1 = (int)(ev >> 16);
                const bool in0 = i0 < rem, in1 = i1 < rem;
                if (i0 != 0xffff)
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(d0 + 8 * k), "l"(lb + (in0 ? i0 : 0)), "r"(in0 ? 4 : 0));
                if (i1 != 0xffff)
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(d0 + 8 * k + 4), "l"(lb + (in1 ? i1 : 0)), "r"(in1 ? 4 : 0));
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        };
        const int nseg = (Tmax + SEG - 1) / SEG;
        stage_seg(0);
        for (int q = 0; q < nseg; q++) {
            if (q + 1 < nseg) stage_seg(q + 1);
            else asm volatile("cp.async.commit_group;" ::: "memory");
            asm volatile("cp.async.wait_group 1;" ::: "memory");     // segment q has landed (every lane reads only its own copies)
            const float2* __restrict__ row = pairs[warp][0] + lane * ROWF2 + SEG * (q % 3);
            uint2* __restrict__ sv = surv + (size_t)q * SEG * SVS;
#pragma unroll 1
            for (int s = 0; s < SEG; s += 2) {
                const uint2 w0 = acs(m, n, row[s]);
                sv[(size_t)s * SVS] = w0;
                const uint2 w1 = acs(n, m, row[s + 1]);
                sv[(size_t)(s + 1) * SVS] = w1;
            }
        }
        asm volatile("cp.async.wait_group 0;" ::: "memory");
#else
        stage(0, 0);
        stage_wait();
        for (int c = 0; c < nch; c++) {
            if (c + 1 < nch) stage(c + 1, (c + 1) & 1);
            const float2* __restrict__ row = pairs[warp][c & 1] + lane * ROWF2;
            uint2* __restrict__ sv = surv + (size_t)c * CS * SVS;
#pragma unroll 1
            for (int s = 0; s < CS; s += C8B_TP_UNROLL) {
#pragma unroll
                for (int u = 0; u < C8B_TP_UNROLL; u += 2) {
                    const uint2 w0 = acs(m, n, row[s + u]);
                    sv[(size_t)(s + u) * SVS] = w0;
                    const uint2 w1 = acs(n, m, row[s + u + 1]);
                    sv[(size_t)(s + u + 1) * SVS] = w1;
                }
            }
            stage_wait();
        }
#endif

        // ---------------- traceback (lib/decode_impl.cc:282-302), final state 0 ----------------
        {
            uint32_t s = 0, acc = 0;
            constexpr int TB = TBK;                                  // decision words per block; the next block is in flight
            uint2 wa[TB], wb[TB];
            auto fetch = [&](uint2 (&w)[TB], int tb) {
                if (tb >= 0 && tb + TB <= T) {                       // whole block inside the frame: plain loads
                    const uint2* __restrict__ p = surv + (size_t)tb * SVS;
#pragma unroll
                    for (int k = 0; k < TB; k++) w[k] = p[(size_t)k * SVS];
                    return;
                }
#pragma unroll
                for (int k = 0; k < TB; k++) w[k] = (tb >= 0 && tb + k < T) ? surv[(size_t)(tb + k) * SVS] : make_uint2(0u, 0u);
            };
            auto walk = [&](const uint2 (&w)[TB], int tb) {
                if (TB == 32 && tb + TB <= T) {
                    // the whole block lies inside the frame: no per-step test, and state + decoded bits share one register --
                    // S <- 2 S + decision never drops a bit, so what leaves the 6 state bits IS the decoded sequence (bit 5 of
                    // the state entered at step t is that step's input bit); it is collected twice per block
                    uint32_t S = s;
#pragma unroll
                    for (int k = TB - 1; k >= 0; k--) {
                        const uint32_t x = (S & 32u) ? w[k].y : w[k].x;
                        S = (S << 1) | ((x >> (S & 31u)) & 1u);
                        if (k == 16) { acc = (S >> 6) << 16; S &= 63u; }
                    }
                    words[(size_t)(tb >> 5) * TPB] = acc | (S >> 6);
                    acc = 0;
                    s = S & 63u;
                    return;
                }
#pragma unroll
                for (int k = TB - 1; k >= 0; k--) {
                    const int t = tb + k;
                    if (t < T) {
                        acc = (acc << 1) | (s >> 5);                 // decoded bit of step t = input bit of the state entered
                        const uint32_t d = ((s & 32u ? w[k].y : w[k].x) >> (s & 31u)) & 1u;
                        s = ((s & 31u) << 1) | d;
                        if ((t & 31) == 0) { words[(size_t)(t >> 5) * TPB] = acc; acc = 0; }
                    }
                }
            };
            int tb = ((Tmax - 1) / TB) * TB;
            fetch(wa, tb);
            for (; tb >= 0; tb -= 2 * TB) {
                fetch(wb, tb - TB);
                walk(wa, tb);
                fetch(wa, tb - 2 * TB);
                if (tb - TB >= 0) walk(wb, tb - TB);
            }
        }
        if (T <= 0) continue;                                        // (lanes without a frame are done; no warp-level sync below)
        const int nwords = (T + 31) >> 5;
        if (scram != nullptr) {
            uint8_t* so = scram + (size_t)f * scramStride;
            for (int i = 0; i < T && i < scramStride; i++) so[i] = (uint8_t)((words[(size_t)(i >> 5) * TPB] >> (i & 31)) & 1u);
        }

        // ---------------- descramble (lib/decode_impl.cc:304-323) ----------------
        {
            const uint32_t w0 = words[0];
            int st = 0;
#pragma unroll
            for (int i = 0; i < 7; i++) st |= (int)((w0 >> i) & 1u) << (6 - i);
            uint32_t q[6];
#pragma unroll
            for (int wq = 0; wq < 5; wq++) {
                uint32_t v = 0;
                for (int b = 0; b < 32; b++) {
                    const int fb = ((st >> 6) ^ (st >> 3)) & 1;
                    st = ((st << 1) & 0x7e) | fb;
                    v |= (uint32_t)fb << b;
                }
                q[wq] = v;
            }
            q[5] = 0;
            for (int w = 0; w < nwords; w++) {
                uint32_t v = words[(size_t)w * TPB];
                if (w == 0) v = (v ^ (q[0] << 7)) & ~0x7fu;
                else {
                    const int o = (32 * w - 7) % 127, k = o >> 5;
                    const uint32_t lo = k == 0 ? q[0] : k == 1 ? q[1] : k == 2 ? q[2] : q[3];
                    const uint32_t hi = k == 0 ? q[1] : k == 1 ? q[2] : k == 2 ? q[3] : q[4];
                    v ^= __funnelshift_r(lo, hi, o & 31);
                }
                words[(size_t)w * TPB] = v;
            }
        }

        // ---------------- packetAssemble (lib/decode_impl.cc:325-520) ----------------
        {
            uint8_t* out = pdu + (size_t)f * pduStride;
            const int cap = (int)min(pduStride, (int64_t)0x7fffffff);
            int npdu = 0, w = 0;
            if (fmt == C8B_F_VHT) {
                int procd = 16;
                if (procd < T) {
                    int bp = 2, tl = 0;                              // tl is NOT reset per subframe (:336)
                    while (true) {
                        procd += 32;
                        if (procd > T) break;
                        const int d0 = (int)get_byte(words, bp), d1 = (int)get_byte(words, bp + 1);
                        const int eof = d0 & 1;
                        tl |= ((d0 >> 2) & 1) << 12;
                        tl |= ((d0 >> 3) & 1) << 13;
                        tl |= (d0 >> 4) | (d1 << 4);
                        const int padded = (tl / 4 + ((tl % 4) != 0)) * 4;
                        procd += padded * 8;
                        if (procd > T) break;
                        bp += 4;
                        if (crc32_words(crcTab, words, bp, tl) == 558161692u) {
                            emit_record_t(out, w, cap, npdu, fmt, tl, words, bp, tl, mcs);
                            tl += 4;                                 // :415, carried into the next subframe
                        }
                        bp += padded;
                        if (eof) break;
                    }
                }
            } else if (!ampdu) {
                if (len >= 0 && 16 + 8 * len <= 32 * nwords) {
                    if (crc32_words(crcTab, words, 2, len) == 558161692u) emit_record_t(out, w, cap, npdu, fmt, len, words, 2, len, mcs);
                }
            }
            frames[f].npdu = npdu; frames[f].pdu_bytes = w;
        }
    }
}

}  // namespace

// C8B_TP_RESIDENT: CTAs per SM a launch actually places (default: all C8B_TP_CTAS the registers allow).  With fewer, the dynamic
// shared memory is padded so that no more fit, and the registers / shared memory left over host front-end CTAs of the next chunk.
#ifndef C8B_TP_RESIDENT
#define C8B_TP_RESIDENT C8B_TP_CTAS
#endif
static constexpr size_t tp_smem_need() { return sizeof(float2) * (TPB / 32) * NBUF * 32 * ROWF2 + 1024 + 512; }
static constexpr size_t tp_smem_bytes()
{
    // one more CTA than wanted must not fit: (R + 1) * (bytes + 1 KB reserved per CTA) > 228 KB
    return C8B_TP_RESIDENT < C8B_TP_CTAS && tp_smem_need() <= (size_t)(228 * 1024 / (C8B_TP_RESIDENT + 1) - 1024)
               ? (size_t)(228 * 1024 / (C8B_TP_RESIDENT + 1) - 1024 + 256) : tp_smem_need();
}
static constexpr size_t tp_surv_per_cta() { return (size_t)(C8B_DECODE_T_MAX + CS + 2) * TPB; }              // uint2
static constexpr size_t tp_words_per_cta() { return (size_t)((C8B_DECODE_T_MAX + 63) / 32 + 1) * TPB; }      // uint32
static int tp_grid(int num_sm, int nframes)
{
    const int full = num_sm * C8B_TP_RESIDENT, need = (nframes + TPB - 1) / TPB;
    return need < full ? need : full;
}

// The dynamic shared-memory opt-in (> 48 KB) is a per-DEVICE function attribute: every context sets it on its own device
// at c8b_create (two contexts on two GPUs in one process, block threads created concurrently).
cudaError_t c8b_viterbi_tp_prepare(void)
{
    return cudaFuncSetAttribute(k_viterbi_tp, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tp_smem_bytes());
}

// scratch for a launch over nframes frames: per resident CTA the survivor words and the decoded words
size_t c8b_viterbi_tp_scratch_bytes(int num_sm, int nframes)
{
    return (size_t)tp_grid(num_sm, nframes) * (tp_surv_per_cta() * sizeof(uint2) + tp_words_per_cta() * sizeof(uint32_t));
}

int c8b_viterbi_tp_wave(int num_sm) { return num_sm * C8B_TP_RESIDENT * TPB; }   // frames in one full wave

void c8b_launch_viterbi_tp(const c8b_lut* d_lut, c8b_frame* d_frames, int nframes, const float* d_llr, int64_t nllr, void* d_scratch,
                           int num_sm, uint8_t* d_pdu, int64_t pdu_stride, uint8_t* d_scram, int64_t scram_stride, cudaStream_t st)
{
    if (nframes <= 0) return;
    const int grid = tp_grid(num_sm, nframes);
    uint2* surv = reinterpret_cast<uint2*>(d_scratch);
    uint32_t* words = reinterpret_cast<uint32_t*>(surv + (size_t)grid * tp_surv_per_cta());
    k_viterbi_tp<<<grid, TPB, tp_smem_bytes(), st>>>(d_lut, d_frames, nframes, d_llr, nllr, surv, words, tp_surv_per_cta(), tp_words_per_cta(), d_pdu,
                                                     pdu_stride, d_scram, scram_stride);
}
