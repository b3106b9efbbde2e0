// k_frontend_w.cu -- warp-cooperative versions of the per-frame stages:
//   k_detect_w : one WARP per item  : trigger FSM -> sync -> signal   (same results as k_detect)
//   k_header_w : one WARP per frame : demod header states, 1 antenna  (same results as k_header)
//   k_trig_scan / k_cand_eval / k_cand_accept : few long items with many frames each (a capture, a stream window)
// The thread-per-item kernels in k_frontend.cu are latency-bound (5 KB of local arrays per thread, ~7 warps per
// SM at a 37888-item chunk); here the control flow stays the reference's sequential state machine, executed
// uniformly by all 32 lanes, while every heavy loop is spread over the lanes through a per-warp shared-memory
// workspace: the 240 conjugate products / powers of the LTF autocorrelation, the per-lag normalisation, the CFO
// rotations, the 64-point DFTs (8 lanes each, the same 8x8 scheme as k_demod), the per-tone SIG demodulation and the
// 64-state SIG Viterbi (2 states per lane, decisions by ballot).  Sums whose order matters in the reference (the
// running sums of lib/sync_impl.cc:155-179, the LTF CFO sum, pilot sums) are still formed sequentially, in order.
// DFTs are float32 here (the thread kernels use a double DFT); both stand in for FFTW (unpinned).
#include "common.cuh"
#include "phy_serial.cuh"

namespace {

using namespace c8b;

constexpr int FW = 4;                          // warps (= items / frames) per CTA
constexpr unsigned FULL = 0xffffffffu;

struct __align__(16) Ws {                       // per-warp workspace
    float2 A[256];                              // complex scratch: products / rotated windows / spectra
    float2 B[256];
    union {
        float4 Lg[112];                         // per-lag (msum.re, msum.im, s1, s2): sync only, which runs no DFT
        float2 X[4][72];                        // DFT transposes (row stride 9)
    };
    float F[256];                               // float scratch: powers / ac / soft bits
    float M[2][64];                             // Viterbi metrics
    uint32_t Dec[48][2];                        // Viterbi decisions
};

struct cpx { float x, y; };
__device__ __forceinline__ cpx operator+(cpx a, cpx b) { return { a.x + b.x, a.y + b.y }; }
__device__ __forceinline__ cpx operator-(cpx a, cpx b) { return { a.x - b.x, a.y - b.y }; }
__device__ __forceinline__ cpx cm(cpx a, cpx b) { return { a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x }; }
__device__ __forceinline__ cpx mulmj(cpx a) { return { a.y, -a.x }; }
__device__ __forceinline__ void dft8(cpx* v)
{
    const float r = 0.70710678118654752440f;
    cpx a0 = v[0] + v[4], a1 = v[0] - v[4], a2 = v[2] + v[6], a3 = mulmj(v[2] - v[6]);
    cpx a4 = v[1] + v[5], a5 = v[1] - v[5], a6 = v[3] + v[7], a7 = mulmj(v[3] - v[7]);
    cpx b0 = a0 + a2, b2 = a0 - a2, b1 = a1 + a3, b3 = a1 - a3;
    cpx b4 = a4 + a6, b6 = mulmj(a4 - a6), b5 = a5 + a7, b7 = a5 - a7;
    b5 = { (b5.x + b5.y) * r, (b5.y - b5.x) * r };
    b7 = { (b7.y - b7.x) * r, -(b7.x + b7.y) * r };
    v[0] = b0 + b4; v[4] = b0 - b4; v[1] = b1 + b5; v[5] = b1 - b5;
    v[2] = b2 + b6; v[6] = b2 - b6; v[3] = b3 + b7; v[7] = b3 - b7;
}

// up to 4 independent 64-point forward DFTs per warp: group g = lane/8 transforms buf + 64*g in place (g < nfft)
__device__ __forceinline__ void fft64_groups(const c8b_lut* __restrict__ L, float2* buf, Ws& W, int nfft, int lane)
{
    const int g = lane >> 3, j = lane & 7;
    cpx v[8];
    float2* b = buf + 64 * g;
    __syncwarp();
    if (g < nfft) {
#pragma unroll
        for (int m = 0; m < 8; m++) { const float2 t = b[j + 8 * m]; v[m] = { t.x, t.y }; }
        dft8(v);
#pragma unroll
        for (int k1 = 1; k1 < 8; k1++) { const int t = (j * k1) & 63; v[k1] = cm(v[k1], cpx{ L->twr[t], L->twi[t] }); }
#pragma unroll
        for (int k1 = 0; k1 < 8; k1++) W.X[g][k1 * 9 + j] = make_float2(v[k1].x, v[k1].y);
    }
    __syncwarp();
    if (g < nfft) {
#pragma unroll
        for (int n1 = 0; n1 < 8; n1++) { const float2 t = W.X[g][j * 9 + n1]; v[n1] = { t.x, t.y }; }
        dft8(v);
#pragma unroll
        for (int k2 = 0; k2 < 8; k2++) b[j + 8 * k2] = make_float2(v[k2].x, v[k2].y);
    }
    __syncwarp();
}

__device__ __forceinline__ cf ld(const float2* p) { const float2 t = *p; return mk(t.x, t.y); }
__device__ __forceinline__ float2 st(cf a) { return make_float2(a.re, a.im); }

// SIG-field Viterbi (lib/cloud80211phy.cc:2001-2088), T <= 48: lane k owns the butterfly (2k, 2k+1) -> (k, k+32).
// Returns the decoded bits packed LSB-first (bit t of lo/hi words).
__device__ __forceinline__ uint64_t sig_viterbi_w(const c8b_lut* __restrict__ L, const float* __restrict__ llr, int T, Ws& W, int lane)
{
    const int c = L->bmClass[lane];
    __syncwarp();
    W.M[0][lane] = lane == 0 ? 0.0f : -1000000000000000.0f;
    W.M[0][lane + 32] = -1000000000000000.0f;
    __syncwarp();
    for (int t = 0; t < T; t++) {
        const float t0 = llr[2 * t], t1 = llr[2 * t + 1];
        const float t3 = fadd(t1, t0);
        const float A = c == 0 ? 0.0f : c == 1 ? t1 : c == 2 ? t0 : t3;
        const float Bm = c == 0 ? t3 : c == 1 ? t0 : c == 2 ? t1 : 0.0f;
        const float* __restrict__ pre = W.M[t & 1];
        float* __restrict__ cur = W.M[(t & 1) ^ 1];
        const float e = pre[2 * lane], o = pre[2 * lane + 1];
        const float a0 = fadd(e, A), b0 = fadd(o, Bm), a1 = fadd(e, Bm), b1 = fadd(o, A);
        float v0 = -1000000000000000.0f, v1 = -1000000000000000.0f;
        bool d0 = false, d1 = false;
        if (a0 > v0) v0 = a0;
        if (b0 > v0) { v0 = b0; d0 = true; }
        if (a1 > v1) v1 = a1;
        if (b1 > v1) { v1 = b1; d1 = true; }
        const uint32_t w0 = __ballot_sync(FULL, d0), w1 = __ballot_sync(FULL, d1);
        if (lane == 0) { W.Dec[t][0] = w0; W.Dec[t][1] = w1; }
        cur[lane] = v0; cur[lane + 32] = v1;
        __syncwarp();
    }
    uint64_t bits = 0;
    int s = 0;                                                    // final state 0
    for (int t = T - 1; t >= 0; t--) {
        bits |= (uint64_t)(s >> 5) << t;
        s = ((s & 31) << 1) | (int)((W.Dec[t][s >> 5] >> (s & 31)) & 1u);
    }
    return bits;
}

__device__ __forceinline__ void unpack_bits(uint64_t v, uint8_t* b, int n) { for (int i = 0; i < n; i++) b[i] = (uint8_t)((v >> i) & 1u); }

// ---------------------------------------------------------------------------------------------------
// sync (lib/sync_impl.cc:92-147, :155-179, :181-196), sig = 240 samples from the trigger (global memory)
// ---------------------------------------------------------------------------------------------------
__device__ SyncOut sync_at_w(const cf* __restrict__ sig, cf conjAvg, Ws& W, int lane)
{
    __syncwarp();
    for (int i = lane; i < C8B_SYNC_BUF; i += 32) {
        const cf s = sig[i];
        W.F[i] = abs2(s);
        if (i < 176) W.A[i] = st(cmul(s, cconj(sig[i + 64])));
    }
    __syncwarp();
    {   // the reference's running sums, in its order (uniform across lanes)
        cf msum = mk(0.f, 0.f);
        float s1 = 0.f, s2 = 0.f;
        for (int i = 0; i < 64; i++) { msum = cadd(msum, ld(&W.A[i])); s1 = fadd(s1, W.F[i]); s2 = fadd(s2, W.F[i + 64]); }
        for (int i = 0; i < C8B_SYNC_RES; i++) {
            if (lane == 0) W.Lg[i] = make_float4(msum.re, msum.im, s1, s2);
            msum = csub(msum, ld(&W.A[i])); s1 = fsub(s1, W.F[i]); s2 = fsub(s2, W.F[i + 64]);
            msum = cadd(msum, ld(&W.A[i + 64])); s1 = fadd(s1, W.F[i + 64]); s2 = fadd(s2, W.F[i + 128]);
        }
    }
    __syncwarp();
    float* __restrict__ ac = reinterpret_cast<float*>(W.B);        // 111 normalised correlations
    for (int i = lane; i < C8B_SYNC_RES; i += 32) {
        const float4 g = W.Lg[i];
        ac[i] = fdiv(fdiv(cabsf_(mk(g.x, g.y)), fsqrt(g.z)), fsqrt(g.w));
    }
    __syncwarp();
    // first maximum (:97: best = ac[0], then every strictly larger value in index order), lane-parallel: every lane starts from
    // ac[0] like the serial scan does (a NaN there wins, as in the reference), keeps the first maximum of its own indices,
    // and the warp keeps the larger value -- the smaller index on a tie
    float best = ac[0];
    int bi = 0;
    for (int i = lane; i < C8B_SYNC_RES; i += 32) { const float a = ac[i]; if (a > best) { best = a; bi = i; } }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        const float ob = __shfl_xor_sync(FULL, best, d);
        const int oi = __shfl_xor_sync(FULL, bi, d);
        if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
    }
    SyncOut o; o.ok = 0; o.mIndex = 0; o.rad = o.snr = o.rssi = 0.f;
    if ((double)best > 0.5) {                                     // :99
        const float thr = (float)((double)best * 0.8);
        // shoulders: the nearest index at or below / at or above the maximum whose correlation is under 0.8 of it (:100-110),
        // from four ballots over the 111 values instead of two serial walks
        int l = bi, r = bi;
        uint32_t under[4];
#pragma unroll
        for (int q = 0; q < 4; q++) { const int i = 32 * q + lane; under[q] = __ballot_sync(FULL, i < C8B_SYNC_RES && ac[i] < thr); }
        {
            const int qb = bi >> 5, sb = bi & 31;
            bool found = false;
#pragma unroll
            for (int q = 3; q >= 0; q--) {
                if (q > qb || found) continue;
                const uint32_t m = q == qb ? under[q] & (0xffffffffu >> (31 - sb)) : under[q];
                if (m) { l = 32 * q + 31 - __clz(m); found = true; }
            }
            found = false;
#pragma unroll
            for (int q = 0; q < 4; q++) {
                if (q < qb || found) continue;
                const uint32_t m = q == qb ? under[q] & (0xffffffffu << sb) : under[q];
                if (m) { r = 32 * q + __ffs(m) - 1; found = true; }
            }
        }
        o.ok = 1; o.mIndex = (l + r) / 2;
        const cf* __restrict__ s = sig + o.mIndex;                // ltf_cfo :181-196
        const float radStf = fdiv(atan2f_(conjAvg.im, conjAvg.re), 16.0f);
        const float bestPwr = W.Lg[bi].z;
        __syncwarp();
        for (int i = lane; i < 128; i += 32) W.A[i] = st(cmul(s[i], cis(fmul((float)i, radStf))));
        __syncwarp();
        for (int i = lane; i < 64; i += 32) W.A[128 + i] = st(cmul(ld(&W.A[i]), cconj(ld(&W.A[i + 64]))));
        __syncwarp();
        cf csum = mk(0.f, 0.f);
        for (int i = 0; i < 64; i++) csum = cadd(csum, ld(&W.A[128 + i]));
        const cf c64 = cdivs(csum, 64.0f);
        const float radLtf = fdiv(atan2f_(c64.im, c64.re), 64.0f);
        o.rad = fadd(radStf, radLtf);
        const double maxD = (double)best;
        o.snr = (float)(10.0 * log10(maxD / (1.0 - maxD)));
        o.rssi = fdiv(bestPwr, 64.0f);
    }
    __syncwarp();
    return o;
}

// ---------------------------------------------------------------------------------------------------
// signal, S_DEMOD (lib/signal_impl.cc:108-162): in = samples from the sync index (>= 224); h -> global (64 complex)
// ---------------------------------------------------------------------------------------------------
__device__ int signal_at_w(const c8b_lut* __restrict__ L, const cf* __restrict__ in, float rad, float2* __restrict__ hOut, int* mcs, int* len,
                           int* nsamp, Ws& W, int lane)
{
    __syncwarp();
    for (int k = lane; k < 192; k += 32) {                        // :115-120, windows at 8, 72, 152
        const int w = k >> 6, i = k & 63, off = w == 0 ? C8B_SYM_SHIFT : w == 1 ? C8B_SYM_SHIFT + 64 : C8B_SYM_SHIFT + 144;
        W.A[k] = st(cmul(in[off + i], cis(fmul((float)(i + off), rad))));
    }
    fft64_groups(L, W.A, W, 3, lane);                             // A[0..63] = LTF1, [64..127] = LTF2, [128..191] = L-SIG
    // procLHSigDemodDeint (lib/cloud80211phy.cc:609-627)
    const int pb[4] = { 7, 21, 43, 57 };
    cf hp[4];
#pragma unroll
    for (int q = 0; q < 4; q++) hp[q] = cdivs(cadd(ld(&W.A[pb[q]]), ld(&W.A[64 + pb[q]])), fmul(2.0f, L->ltfL[pb[q]]));
    cf acc = cdiv(ld(&W.A[128 + 7]), hp[0]);
    acc = csub(acc, cdiv(ld(&W.A[128 + 21]), hp[1]));
    acc = cadd(acc, cdiv(ld(&W.A[128 + 43]), hp[2]));
    acc = cadd(acc, cdiv(ld(&W.A[128 + 57]), hp[3]));
    const cf ps = cconj(acc);
    const float pa = cabsf_(ps);
    for (int i = lane; i < 64; i += 32) {
        cf h = mk(0.f, 0.f);
        const int d = L->sigDemap[i];
        const bool pil = i == 7 || i == 21 || i == 43 || i == 57;
        if (d >= 0 || pil) h = cdivs(cadd(ld(&W.A[i]), ld(&W.A[64 + i])), fmul(2.0f, L->ltfL[i]));
        if (d >= 0) W.F[d] = cdivs(cmul(cdiv(ld(&W.A[128 + i]), h), ps), pa).re;
        hOut[i] = st(h);
    }
    __syncwarp();
    const uint64_t bits = sig_viterbi_w(L, W.F, 24, W, lane);
    uint8_t b[24];
    unpack_bits(bits, b, 24);
    if (!lsig_check(b, mcs, len)) return 0;
    const int ndbps = lsig_ndbps(*mcs);
    *nsamp = ((*len * 8 + 22 + ndbps - 1) / ndbps) * 80;          // :128
    return 1;
}

// A complete word that ends below the threshold, holds no run of 21 (counting what the previous word left in nPlateau) and
// meets the trigger idle with no hold-off pending cannot leave a trace (see the short-run rule in walk_word): the FSM is
// in its reset state after it.
__device__ __forceinline__ bool word_traceless(const TrigState& ts, const uint32_t m, const int kmax, const int i0, const int skipUntil,
                                               const bool muted)
{
    if (ts.fPlateau || kmax < 32 || (m >> 31) || !(i0 >= skipUntil || muted)) return false;
    uint32_t x = m;
    x &= x >> 1; x &= x >> 2; x &= x >> 4; x &= x >> 8; x &= x >> 5;         // a bit survives iff a run of >= 21 ones exists
    return x == 0u && ts.nPlateau + (__ffs(~m) - 1) <= 20;
}

// One bitmap word (samples i0 .. i0 + kmax) of the trigger FSM (lib/trigger_impl.cc:59-117, trig_step in phy_serial.cuh), evaluated
// by the warp without walking every sample.  pv = this lane's preac value.  The word is cut at the only samples where the
// FSM changes course -- the 21st sample of a plateau (count-down armed) and the last sample of the count-down (trigger) --
// which go through trig_step itself; everything between is applied in bulk:
//   * a stretch below the threshold resets the plateau state and advances the count-down,
//   * a short run that ends inside the word with the trigger idle is stepped over (it cannot leave a trace),
//   * a run above the threshold adds to nPlateau / the count-down; the flag 0x02 ("new maximum", the latch of sync's input
//     2) fires at every sample that beats the running maximum, the last of which is the FIRST occurrence of the run's
//     maximum -- one warp max + ballot instead of a walk.
// Events before skipUntil (sync's 111-sample hold-off, lib/sync_impl.cc:141-146) or after a stall are dropped exactly as
// the serial loop drops them.  TRACK (stream windows): safe = latest sample before which the FSM is in its reset state.
// Resumable: returns the index of the next trigger (0x01) the caller has to act on, with k advanced past it, or -1 when the
// word is finished; the caller handles the trigger (it may move skipUntil / set muted) and calls again.
template <bool TRACK>
__device__ __forceinline__ int walk_word(TrigState& ts, const uint32_t m, int& k, const int i0, const int kmax, const float pv, const int lane,
                                         const int skipUntil, const bool muted, int& latch, int& safe)
{
    while (k < kmax) {
        const int i = i0 + k;
        if (TRACK && ts.nPlateau == 0 && ts.fPlateau == 0 && i >= skipUntil) safe = i;
        const uint32_t rest = m >> k;
        if (!(rest & 1u)) {                                                 // below the threshold: samples k .. k + gap - 1
            const int gap = rest ? __ffs(rest) - 1 : kmax - k;
            ts.nPlateau = 0; ts.fPlateauEnd = 0; ts.conjAc = 0.0f;
            if (ts.fPlateau) {
                if (ts.countDown <= gap) {                                  // the count-down ends in here: 0x01 at its last sample
                    k += ts.countDown;
                    ts.countDown = 0; ts.fPlateau = 0;
                    const int t = i0 + k - 1;
                    if (t >= skipUntil && !muted) return t;
                    continue;
                }
                ts.countDown -= gap;
            }
            k += gap;
            continue;
        }
        const uint32_t inv = ~rest;
        int P = min(inv ? __ffs(inv) - 1 : 32, kmax - k);                   // run of samples above the threshold
        // A short run that ends inside the word while the trigger is idle leaves nothing behind: the sample after it resets
        // nPlateau / conjAc, and its 0x02 flags only move a latch that the first sample of the next plateau overwrites
        // before any trigger can read it (no hold-off is pending, so that flag is honoured).
        if (!ts.fPlateau && k + P < kmax && ts.nPlateau + P <= 20 && (i >= skipUntil || muted)) { k += P; continue; }
        if (ts.fPlateau) P = min(P, ts.countDown - 1);                      // ... up to the sample that ends the count-down
        else if (ts.fPlateauEnd == 0) P = min(P, max(0, 20 - ts.nPlateau)); // ... up to the sample that arms it (nPlateau 21)
        if (P > 0) {
            const float v = (lane >= k && lane < k + P) ? pv : -1.0f;       // preac >= 0
            float mx = v;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(FULL, mx, o));
            if (mx > ts.conjAc) {                                           // :82-87, the last 0x02 of the run
                const int first = i0 + __ffs(__ballot_sync(FULL, v == mx)) - 1;
                ts.conjAc = mx;
                if (first >= skipUntil && !muted) latch = first;
            }
            ts.nPlateau += P;
            if (ts.fPlateau) ts.countDown -= P;
            k += P;
            continue;
        }
        const uint8_t fl = trig_step(ts, __shfl_sync(FULL, pv, k));         // the sample that arms or ends the count-down
        k++;
        if (fl == 0 || i < skipUntil || muted) continue;
        if (fl & 0x01) return i;
        if (fl & 0x02) latch = i;
    }
    return -1;
}

__global__ void __launch_bounds__(FW * 32, 6)
k_detect_w(const c8b_lut* __restrict__ lut, const float2* __restrict__ iq, const int64_t* __restrict__ off,
           const int32_t* __restrict__ len, int nitems, int itemBase, int maxf, int64_t outBase, const float* __restrict__ preacAll,
           const uint32_t* __restrict__ maskAll, int maskStride, c8b_frame* __restrict__ frames, float2* __restrict__ chan)
{
    __shared__ Ws ws[FW];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int it = blockIdx.x * FW + warp;
    if (it >= nitems) return;
    Ws& W = ws[warp];
    const cf* __restrict__ x = reinterpret_cast<const cf*>(iq + off[it]);
    const float* __restrict__ preac = preacAll + (off[it] - outBase);
    const uint32_t* __restrict__ mask = maskAll ? maskAll + (size_t)it * maskStride : nullptr;
    const int n = len[it];
    c8b_frame* __restrict__ f = frames + (size_t)it * maxf;
    float2* __restrict__ h = chan + (size_t)it * maxf * 64;
    if (lane == 0) for (int k = 0; k < maxf; k++) frame_clear(f + k, itemBase + it, C8B_ST_EMPTY);
    for (int k = lane; k < maxf * 64; k += 32) h[k] = make_float2(0.f, 0.f);
    __syncwarp();

    // the blocks' state machines, evaluated uniformly by the warp (see detect_item in phy_serial.cuh)
    TrigState ts;
    trig_reset(ts);
    int latch = -1, skipUntil = 0, nTrig = 0, nEv = 0, nLsigFail = 0, pos = 0, nf = 0;
    bool syncStalled = false, sigStalled = false, done = false;
    // The scan goes bitmap word by bitmap word (32 samples).  A complete word without a sample above the threshold is
    // not walked: with the trigger idle the FSM stays in its reset state (lib/trigger_impl.cc:95-100) -- all such words up
    // to the next flagged one are skipped with one ballot over 32 prefetched words; during the 80-sample count-down
    // (:101-109) the counter is simply advanced.  Flagged words go through walk_word with a lane-held copy of their 32
    // preac values.
    int blkBase = -(1 << 30);                                        // 32 bitmap words [blkBase, blkBase+32) summarised in nz
    uint32_t nz = 0;                                              // bit k: word blkBase+k must be walked
    const int nwords = (n + 31) >> 5;
    for (int w = 0; w < nwords && !done;) {
        if (mask) {
            if (w < blkBase || w >= blkBase + 32) {
                blkBase = w;
                const int ww = w + lane;
                const uint32_t mv = (ww * 32 + 32 <= n) ? mask[ww] : 0xffffffffu;      // a partial last word is always walked
                nz = __ballot_sync(FULL, mv != 0u);
            }
            const uint32_t rem = nz >> (w - blkBase);
            if (!(rem & 1u)) {
                if (ts.fPlateau == 0) {
                    w += rem ? __ffs(rem) - 1 : 32 - (w - blkBase);
                    ts.nPlateau = 0; ts.fPlateauEnd = 0; ts.conjAc = 0.0f;
                    continue;
                }
                if (ts.countDown > 32) {                          // 32 sub-threshold samples of the count-down
                    ts.countDown -= 32;
                    ts.nPlateau = 0; ts.fPlateauEnd = 0; ts.conjAc = 0.0f;
                    w++;
                    continue;
                }
            }
        }
        const int i0 = w * 32;
        const float pv = (i0 + lane < n) ? preac[i0 + lane] : 0.0f;
        const int kmax = min(32, n - i0);
        auto onTrig = [&](int i) {
            nTrig++;
            if (n - i < C8B_SYNC_BUF) { syncStalled = true; return; }
            const cf cj = latch >= 0 ? presiso_conj_at(x, latch) : mk(0.f, 0.f);
            const SyncOut so = sync_at_w(x + i, cj, W, lane);
            skipUntil = i + C8B_SYNC_RES;
            if (!so.ok) return;
            nEv++;
            const int idx = i + so.mIndex;
            if (sigStalled || idx < pos) return;
            if (n - idx < 224) { sigStalled = true; return; }
            int mcs = 0, ln = 0, nsamp = 0;
            if (!signal_at_w(lut, x + idx, so.rad, h + nf * 64, &mcs, &ln, &nsamp, W, lane)) {
                for (int q = lane; q < 64; q += 32) h[nf * 64 + q] = make_float2(0.f, 0.f);
                nLsigFail++; pos = idx + 80; return;
            }
            c8b_frame* fk = f + nf;
            nf++;
            pos = idx + 224 + nsamp;
            const int status = pos > n ? C8B_ST_TRUNC : C8B_ST_OK;
            if (lane == 0) {
                fk->trig_idx = i; fk->sync_idx = idx; fk->rad = so.rad; fk->snr = so.snr; fk->rssi = so.rssi;
                fk->cfo_hz = fmul(so.rad, 3183098.8618379068f);
                fk->l_mcs = mcs; fk->l_len = ln; fk->nsamp = nsamp; fk->status = status;
            }
            if (pos > n || nf >= maxf) done = true;
        };
        const uint32_t above = __ballot_sync(FULL, lane < kmax && pv > 0.3f);      // lib/trigger_impl.cc:79
        if (word_traceless(ts, above, kmax, i0, skipUntil, syncStalled)) { ts.nPlateau = 0; ts.fPlateauEnd = 0; ts.conjAc = 0.0f; w++; continue; }
        int unused = 0;
        for (int k = 0, t; !done && (t = walk_word<false>(ts, above, k, i0, kmax, pv, lane, skipUntil, syncStalled, latch, unused)) >= 0;) onTrig(t);
        w++;
    }
    if (nf == 0 && lane == 0) f->status = nTrig == 0 ? C8B_ST_NO_TRIGGER : nEv == 0 ? C8B_ST_SYNC : (nLsigFail ? C8B_ST_LSIG : C8B_ST_TRUNC);
}

// ---------------------------------------------------------------------------------------------------
// Many frames per item (a long capture, a window of a live stream): the serial part of detection is only the trigger
// FSM and the accept rules; sync + signal of every trigger are independent of each other.  Three passes:
//   k_trig_scan  : one warp per item -- the FSM scan of k_detect_w without the per-trigger work; emits the triggers that
//                  survive the sync hold-off as candidates (index, latch, safe point before their plateau)
//   k_cand_eval  : one warp per candidate -- sync_at_w, then signal_at_w when the LTF correlation passes
//   k_cand_accept: one thread per item -- S_COPY swallow rule, stalls, frame records, stream restart point
// Same results as the one-thread routine detect_item (phy_serial.cuh; frontend_mode 1 -- the tests run both); used when few
// items may hold many frames each.
// ---------------------------------------------------------------------------------------------------
struct Cand {
    int32_t trig, latch, safe;      // trigger index, argmax of preac in its plateau, restart point before the plateau
    int32_t stall;                  // 1: fewer than 240 samples after the trigger (sync_impl.cc:94)
    int32_t ok, idx, sig;           // sync passed, sync index, signal: 0 not run (stall), 1 L-SIG ok, 2 L-SIG failed
    int32_t mcs, len, nsamp;
    float rad, snr, rssi;
    float2 chan[64];
};
struct CandHead { int32_t n, overflow, safeEnd, nTrig; };

// The scan itself is cut into SEGMENTS that run in parallel, one warp each.  A cut may sit at any bitmap word that follows
// C8B_CUT_WORDS all-zero words (192 samples below the threshold): the FSM is then in its reset state (nPlateau, conjAc and
// fPlateauEnd clear at the first sample below the threshold; a count-down armed by an earlier plateau has run out after 80
// samples), and no sync hold-off is pending (a trigger fires 80 samples after the sample that armed it and holds sync for
// 111: 191 in all) -- exactly the state the scan of the whole item has there, so the segments' triggers, latches and
// restart points are the ones of one serial pass.  Nominal cuts every C8B_SEG_WORDS words are moved forward to the first
// such word (a segment with no quiet stretch simply runs on into the next); between frames of a real capture the channel
// is quiet for longer than 192 samples (SIFS is 320), so a 1 M-sample window splits into ~128 segments.
constexpr int C8B_CUT_WORDS = 6;               // all-zero bitmap words in front of a cut (>= 191 samples)
constexpr int C8B_SEG_WORDS = 256;             // nominal segment: 8192 samples
constexpr int C8B_SEG_CAP = 96;                // candidates per segment (a hold-off of 111 samples: <= 74 per 8192)

struct SegTab { int32_t nseg; int32_t start[1]; };             // start[0 .. nseg], start[nseg] = item length (variable length)
__host__ __device__ inline size_t segtab_bytes(int nsegMax) { return ((size_t)(nsegMax + 2) * sizeof(int32_t) + 15) & ~(size_t)15; }

__global__ void __launch_bounds__(256)
k_seg_cuts(const int32_t* __restrict__ len, int nitems, const uint32_t* __restrict__ maskAll, int maskStride, const c8b_scan* __restrict__ scans,
           uint8_t* __restrict__ tabs, int nsegMax, int segWords)
{
    const int it = blockIdx.x;
    if (it >= nitems) return;
    SegTab* tb = reinterpret_cast<SegTab*>(tabs + (size_t)it * segtab_bytes(nsegMax));
    const uint32_t* __restrict__ mask = maskAll + (size_t)it * maskStride;
    const int n = len[it], from = scans ? scans[it].from : 0;
    const int nwords = n >> 5;                                  // whole words only: the partial last word stays in the last segment
    const int w0 = from >> 5;
    __shared__ int32_t cut[1024];
    const int nnom = min(nsegMax - 1, 1023);
    for (int k = threadIdx.x + 1; k <= nnom; k += blockDim.x) {
        const long long wk = (long long)w0 + (long long)k * segWords;
        int found = -1;
        if (wk < nwords) {
            int zeros = 0;
            const int lo = (int)wk - C8B_CUT_WORDS, hi = min(nwords, (int)wk + segWords);
            for (int w = max(lo, w0 + 1); w < hi; w++) {       // first w >= wk with words w-6 .. w-1 all zero
                if (w >= (int)wk && zeros >= C8B_CUT_WORDS) { found = w; break; }
                zeros = mask[w] == 0u ? zeros + 1 : 0;
            }
        }
        cut[k] = found;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int ns = 0;
        tb->start[ns++] = from;
        for (int k = 1; k <= nnom; k++) if (cut[k] > 0) tb->start[ns++] = cut[k] * 32;
        tb->start[ns] = n;
        tb->nseg = ns;
    }
}

__global__ void __launch_bounds__(32)
k_trig_scan(const int32_t* __restrict__ len, int nitems, int64_t outBase, const int64_t* __restrict__ off, const float* __restrict__ preacAll,
            const uint32_t* __restrict__ maskAll, int maskStride, const uint8_t* __restrict__ tabs, int nsegMax, Cand* __restrict__ cands,
            CandHead* __restrict__ heads)
{
    const int lane = threadIdx.x, it = blockIdx.x / nsegMax, sg = blockIdx.x % nsegMax;
    if (it >= nitems) return;
    const SegTab* tb = reinterpret_cast<const SegTab*>(tabs + (size_t)it * segtab_bytes(nsegMax));
    if (sg >= tb->nseg) return;
    const float* __restrict__ preac = preacAll + (off[it] - outBase);
    const uint32_t* __restrict__ mask = maskAll + (size_t)it * maskStride;
    const int n = len[it];
    const int maxCand = C8B_SEG_CAP;
    Cand* __restrict__ cd = cands + ((size_t)it * nsegMax + sg) * C8B_SEG_CAP;
    const int from = tb->start[sg], segEnd = tb->start[sg + 1];
    TrigState ts;
    trig_reset(ts);
    int latch = -1, skipUntil = 0, nc = 0, nTrig = 0, safe = from, overflow = 0;
    bool done = false;
    // This warp is alone with its item, so nothing hides a dependent load: the bitmap words of 4 x 32 x 32 samples are fetched
    // at once (four per lane) and the preac values of the flagged words among them are staged in shared memory in one go.
    constexpr int SB = 4;                                            // sub-blocks of 32 words per fetch
    __shared__ float pvb[SB * 1024];
    int supBase = -(1 << 30);
    uint32_t nzs[SB] = { 0, 0, 0, 0 }, mvs[SB] = { 0, 0, 0, 0 };
    const int nwords = (segEnd + 31) >> 5;                       // (cuts are word aligned; the last segment ends with the item)
    for (int w = from >> 5; w < nwords && !done;) {
        const int kfirst = w == (from >> 5) ? (from & 31) : 0;
        int jb = 0, blkBase = 0;
        uint32_t nz = 0, mvLane = 0;
        if (kfirst == 0) {
            if (w < supBase || w >= supBase + 32 * SB) {
                supBase = w;
                __syncwarp();
#pragma unroll
                for (int j = 0; j < SB; j++) {
                    const int ww = supBase + 32 * j + lane;
                    const bool part = ww * 32 < n && ww * 32 + 32 > n;        // the partial last word is always walked
                    mvs[j] = ww * 32 < n ? mask[ww] : 0u;
                    nzs[j] = __ballot_sync(FULL, part || mvs[j] != 0u);
                }
#pragma unroll
                for (int j = 0; j < SB; j++) {
#pragma unroll 8
                    for (int r = 0; r < 32; r++)
                        if ((nzs[j] >> r) & 1u) {
                            const int idx = (supBase + 32 * j + r) * 32 + lane;
                            pvb[(32 * j + r) * 32 + lane] = idx < n ? preac[idx] : 0.0f;
                        }
                }
                __syncwarp();
            }
            jb = (w - supBase) >> 5;
            blkBase = supBase + 32 * jb;
            nz = jb == 0 ? nzs[0] : jb == 1 ? nzs[1] : jb == 2 ? nzs[2] : nzs[3];
            mvLane = jb == 0 ? mvs[0] : jb == 1 ? mvs[1] : jb == 2 ? mvs[2] : mvs[3];
            const uint32_t rem = nz >> (w - blkBase);
            if (!(rem & 1u)) {
                if (ts.fPlateau == 0) {
                    w += rem ? __ffs(rem) - 1 : 32 - (w - blkBase);
                    ts.nPlateau = 0; ts.fPlateauEnd = 0; ts.conjAc = 0.0f;
                    if (w * 32 >= skipUntil) safe = w * 32;
                    continue;
                }
                if (ts.countDown > 32) {
                    ts.countDown -= 32;
                    ts.nPlateau = 0; ts.fPlateauEnd = 0; ts.conjAc = 0.0f;
                    w++;
                    continue;
                }
            }
        }
        const int i0 = w * 32;
        const int kmax = min(32, n - i0);
        float pv;
        uint32_t above;                                               // bit k: sample i0 + k is above the threshold (lib/trigger_impl.cc:79)
        if (kfirst == 0) {
            const int r = w - blkBase;
            pv = ((nz >> r) & 1u) ? pvb[(32 * jb + r) * 32 + lane] : 0.0f;
            above = __shfl_sync(FULL, mvLane, r);                     // k_presiso's bitmap is this very comparison
        } else {                                                      // a window that starts the FSM inside a word
            pv = (i0 + lane < n) ? preac[i0 + lane] : 0.0f;
            above = __ballot_sync(FULL, lane < kmax && pv > 0.3f);
        }
        auto onTrig = [&](int i) {
            nTrig++;
            if (nc >= maxCand) { overflow = 1; done = true; return; }
            const int stall = n - i < C8B_SYNC_BUF;
            if (lane == 0) { cd[nc].trig = i; cd[nc].latch = latch; cd[nc].safe = safe; cd[nc].stall = stall; cd[nc].ok = 0; cd[nc].sig = 0; }
            nc++;
            if (stall) { done = true; return; }                   // live: wait for more samples; batch: nothing after it is looked at
            skipUntil = i + C8B_SYNC_RES;
        };
        if (kfirst == 0 && word_traceless(ts, above, kmax, i0, skipUntil, false)) { ts.nPlateau = 0; ts.fPlateauEnd = 0; ts.conjAc = 0.0f; w++; continue; }
        for (int k = kfirst, t; !done && (t = walk_word<true>(ts, above, k, i0, kmax, pv, lane, skipUntil, false, latch, safe)) >= 0;) onTrig(t);
        w++;
    }
    int safeEnd = -1;                                             // scan ran to the end in the reset state: everything is decided
    if (!done && ts.nPlateau == 0 && ts.fPlateau == 0 && segEnd >= skipUntil) safeEnd = segEnd;
    if (lane == 0) {
        CandHead* hd = heads + (size_t)it * nsegMax + sg;
        hd->n = nc; hd->overflow = overflow; hd->safeEnd = safeEnd >= 0 ? safeEnd : safe; hd->nTrig = nTrig;
    }
}

__global__ void __launch_bounds__(FW * 32)
k_cand_eval(const c8b_lut* __restrict__ lut, const float2* __restrict__ iq, const int64_t* __restrict__ off, const int32_t* __restrict__ len,
            int nitems, Cand* __restrict__ cands, const CandHead* __restrict__ heads, const uint8_t* __restrict__ tabs, int nsegMax)
{
    __shared__ Ws ws[FW];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t gid = (int64_t)blockIdx.x * FW + warp;
    const int64_t slot = gid / C8B_SEG_CAP;                        // (item, segment)
    const int it = (int)(slot / nsegMax), sg = (int)(slot % nsegMax), c = (int)(gid % C8B_SEG_CAP);
    if (it >= nitems) return;
    if (sg >= reinterpret_cast<const SegTab*>(tabs + (size_t)it * segtab_bytes(nsegMax))->nseg || c >= heads[slot].n) return;
    Ws& W = ws[warp];
    Cand* __restrict__ cd = cands + (size_t)slot * C8B_SEG_CAP + c;
    if (cd->stall) return;
    const cf* __restrict__ x = reinterpret_cast<const cf*>(iq + off[it]);
    const int n = len[it], i = cd->trig, latch = cd->latch;
    const cf cj = latch >= 0 ? presiso_conj_at(x, latch) : mk(0.f, 0.f);
    const SyncOut so = sync_at_w(x + i, cj, W, lane);
    int sig = 0, mcs = 0, ln = 0, nsamp = 0;
    const int idx = i + so.mIndex;
    if (so.ok && n - idx >= 224) {
        __syncwarp();
        sig = signal_at_w(lut, x + idx, so.rad, cd->chan, &mcs, &ln, &nsamp, W, lane) ? 1 : 2;
    }
    if (lane == 0) {
        cd->ok = so.ok; cd->idx = idx; cd->rad = so.rad; cd->snr = so.snr; cd->rssi = so.rssi;
        cd->sig = sig; cd->mcs = mcs; cd->len = ln; cd->nsamp = nsamp;
    }
}

__global__ void __launch_bounds__(32)
k_cand_accept(const int32_t* __restrict__ len, int nitems, int itemBase, int maxf, const Cand* __restrict__ cands,
              const CandHead* __restrict__ heads, const uint8_t* __restrict__ tabs, int nsegMax, c8b_frame* __restrict__ frames,
              float2* __restrict__ chan, c8b_scan* __restrict__ scans)
{
    const int it = blockIdx.x, lane = threadIdx.x;                // one warp per item: the rules run uniformly, the copies by lane
    if (it >= nitems) return;
    const int n = len[it];
    const int nseg = reinterpret_cast<const SegTab*>(tabs + (size_t)it * segtab_bytes(nsegMax))->nseg;
    c8b_frame* __restrict__ f = frames + (size_t)it * maxf;
    float2* __restrict__ h = chan + (size_t)it * maxf * 64;
    c8b_scan* sc = scans ? scans + it : nullptr;
    const bool live = sc && !sc->flush;
    for (int k = lane; k < maxf; k += 32) frame_clear(f + k, itemBase + it, C8B_ST_EMPTY);
    for (int k = lane; k < maxf * 64; k += 32) h[k] = make_float2(0.f, 0.f);
    __syncwarp();
    int pos = sc ? sc->pos0 : 0, nf = 0, nEv = 0, nLsigFail = 0;
    int safe = sc ? sc->from : 0, posS = pos, nfS = 0, stalled = 0;
    bool syncStalled = false, sigStalled = false, done = false;
    int nTrig = 0, nCand = 0;
    CandHead hd = heads[(size_t)it * nsegMax];
    for (int sg = 0; sg < nseg && !done; sg++) {                  // the segments in stream order: one candidate list
    hd = heads[(size_t)it * nsegMax + sg];
    const Cand* __restrict__ cd = cands + ((size_t)it * nsegMax + sg) * C8B_SEG_CAP;
    nTrig += hd.nTrig;
    for (int c = 0; c < hd.n && !done; c++) {
        const Cand& q = cd[c];
        if (q.safe > safe || nCand == 0) { safe = q.safe; posS = pos; nfS = nf; }     // the restart point before this trigger's plateau
        nCand++;
        if (syncStalled) continue;
        if (q.stall) {
            if (live) { stalled = 1; done = true; } else syncStalled = true;
            continue;
        }
        if (!q.ok) continue;
        nEv++;
        if (sigStalled || q.idx < pos) continue;                  // swallowed by S_COPY / skipped 80
        if (n - q.idx < 224) {
            if (live) { stalled = 1; done = true; } else sigStalled = true;
            continue;
        }
        if (q.sig != 1) { nLsigFail++; pos = q.idx + 80; continue; }
        c8b_frame* fk = f + nf;
        for (int k = lane; k < 64; k += 32) h[nf * 64 + k] = q.chan[k];
        nf++;
        pos = q.idx + 224 + q.nsamp;
        if (lane == 0) {
            fk->trig_idx = q.trig; fk->sync_idx = q.idx; fk->rad = q.rad; fk->snr = q.snr; fk->rssi = q.rssi;
            fk->cfo_hz = fmul(q.rad, 3183098.8618379068f);
            fk->l_mcs = q.mcs; fk->l_len = q.len; fk->nsamp = q.nsamp;
            fk->status = pos > n ? C8B_ST_TRUNC : C8B_ST_OK;
        }
        if (pos > n) { done = true; stalled = 1; }
        if (nf >= maxf) { done = true; if (!stalled) stalled = 2; }
    }
    if (!done && hd.overflow) { done = true; stalled = 2; }       // candidate records of this segment used up: its scan stopped there
    }
    if (!done && !syncStalled) {
        // the scan's own end state: the restart point the last segment reached after its last candidate
        if (hd.safeEnd > safe || nCand == 0) { safe = hd.safeEnd; posS = pos; nfS = nf; }
    }
    if (lane == 0) {
        if (sc) { sc->safe = safe; sc->pos = posS; sc->nf = nfS; sc->stalled = stalled; }
        if (nf == 0) f->status = nTrig == 0 ? C8B_ST_NO_TRIGGER : nEv == 0 ? C8B_ST_SYNC : (nLsigFail ? C8B_ST_LSIG : C8B_ST_TRUNC);
    }
}

// ---------------------------------------------------------------------------------------------------
// demod header states, one antenna (lib/demod_impl.cc:72-277, :344-505): warp version of demod_header
// ---------------------------------------------------------------------------------------------------
struct RotW {
    const cf* x; float rad; int nsamp;
    __device__ __forceinline__ cf at(int k) const
    {
        if (k >= nsamp) return mk(0.f, 0.f);
        return cmul(x[k], cis(fmul((float)(k + 224), rad)));
    }
};

__global__ void __launch_bounds__(FW * 32, 6)
k_header_w(const c8b_lut* __restrict__ lut, const float2* __restrict__ iq, const int64_t* __restrict__ off, int nslots, int maxf,
           int mupos, c8b_frame* __restrict__ frames, const float2* __restrict__ chan, float2* __restrict__ hinvAll, int64_t llrStride,
           float* __restrict__ llrAll)
{
    __shared__ Ws ws[FW];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int sl = blockIdx.x * FW + warp;
    if (sl >= nslots) return;
    Ws& W = ws[warp];
    c8b_frame* __restrict__ fr = frames + sl;
    float2* __restrict__ hinv = hinvAll + (size_t)sl * 64;
    if (lane == 0) fr->llr_off = (int64_t)sl * llrStride;
    if (fr->status != C8B_ST_OK) return;
    const c8b_lut* L = lut;
    RotW rot;
    rot.x = reinterpret_cast<const cf*>(iq + off[sl / maxf]) + fr->sync_idx + 224;
    rot.rad = fr->rad; rot.nsamp = fr->nsamp;
    const int nsamp = fr->nsamp, lmcs = fr->l_mcs, llen = fr->l_len;
    const float2* __restrict__ hl = chan + (size_t)sl * 64;

    Mod m;
    m.format = m.sumu = m.ampdu = m.nSym = m.nSymSamp = m.nSD = m.nSP = m.nSS = m.nLTF = 0;
    m.mcs = m.len = m.mod = m.cr = m.nBPSCS = m.nDBPS = m.nCBPS = m.nCBPSS = 0;
    const int nsig = nsamp + 320;
    int pos = 0, trellis = 0, status = C8B_ST_OK;
    float sssnr = 0.f;
    bool legacy = lmcs > 0, haveNL = false;
    // window k (0..3) of 64 rotated samples starting at stream index `start` -> W.A[64*k ..]
    auto win = [&](int k, int start) { for (int i = lane; i < 64; i += 32) W.A[64 * k + i] = st(rot.at(start + C8B_SYM_SHIFT + i)); };
    float2* __restrict__ HNL = W.B;                                // [64] non-legacy channel (W.B[0..63])
    float* __restrict__ llrht = W.F, *llrvht = W.F + 96;           // 96 + 96 soft bits

    if (!legacy) {                                                // DEMOD_S_FORMAT :106-148
        if (nsig < 160) status = C8B_ST_TRUNC;
        else {
            __syncwarp();
            win(0, 0); win(1, 80);
            fft64_groups(L, W.A, W, 2, lane);
            {   // procNLSigDemodDeint (lib/cloud80211phy.cc:629-648)
                cf a1 = cdiv(ld(&W.A[7]), ld((const float2*)&hl[7])), a2 = cdiv(ld(&W.A[64 + 7]), ld(&hl[7]));
                a1 = csub(a1, cdiv(ld(&W.A[21]), ld(&hl[21]))); a2 = csub(a2, cdiv(ld(&W.A[64 + 21]), ld(&hl[21])));
                a1 = cadd(a1, cdiv(ld(&W.A[43]), ld(&hl[43]))); a2 = cadd(a2, cdiv(ld(&W.A[64 + 43]), ld(&hl[43])));
                a1 = cadd(a1, cdiv(ld(&W.A[57]), ld(&hl[57]))); a2 = cadd(a2, cdiv(ld(&W.A[64 + 57]), ld(&hl[57])));
                const cf p1 = cconj(a1), p2 = cconj(a2);
                const float m1 = cabsf_(p1), m2 = cabsf_(p2);
                for (int i = lane; i < 64; i += 32) {
                    const int d = L->sigDemap[i];
                    if (d < 0) continue;
                    const cf h = ld(&hl[i]);
                    const cf q1 = cdivs(cmul(cdiv(ld(&W.A[i]), h), p1), m1);
                    const cf q2 = cdivs(cmul(cdiv(ld(&W.A[64 + i]), h), p2), m2);
                    llrht[d] = q1.im; llrht[d + 48] = q2.im;
                    llrvht[d] = q1.re; llrvht[d + 48] = q2.im;
                }
            }
            __syncwarp();
            uint8_t vb[48];
            unpack_bits(sig_viterbi_w(L, llrvht, 48, W, lane), vb, 48);
            if (check_vhta(vb)) {                                 // DEMOD_S_VHT :150-178
                parse_vhta(vb, &m);
                pos = 160;
                const int need = 80 + m.nLTF * 80 + 80;
                if (nsig - pos < need) status = C8B_ST_TRUNC;
                else {
                    __syncwarp();
                    if (m.sumu) {                                 // nonLegacyChanEstimate :344-411
                        win(0, pos + 80); win(1, pos + 160);
                        fft64_groups(L, W.A, W, 2, lane);
                        for (int i = lane; i < 64; i += 32) {
                            cf h = mk(0.f, 0.f);
                            if (!nl_null(i)) {
                                if (mupos == 0) h = cdivs(csub(ld(&W.A[i]), ld(&W.A[64 + i])), fmul(L->ltfNL[i], 2.0f));
                                else h = cdivs(cadd(cdivs(ld(&W.A[i]), L->ltfNL[i]), cdivs(ld(&W.A[64 + i]), L->ltfNL22[i])), 2.0f);
                            }
                            HNL[i] = st(h);
                        }
                    } else {
                        win(0, pos + 80);
                        fft64_groups(L, W.A, W, 1, lane);
                        for (int i = lane; i < 64; i += 32) HNL[i] = nl_null(i) ? make_float2(0.f, 0.f) : st(cdivs(ld(&W.A[i]), L->ltfNL[i]));
                    }
                    haveNL = true;
                    __syncwarp();
                    // vhtSigBDemod :449-505
                    win(0, pos + 80 + m.nLTF * 80);
                    fft64_groups(L, W.A, W, 1, lane);
                    for (int i = lane; i < 64; i += 32) if (!nl_null(i)) W.A[64 + i] = st(cdiv(ld(&W.A[i]), ld(&HNL[i])));   // sig1
                    __syncwarp();
                    const cf ps = cconj(cadd(cadd(csub(ld(&W.A[64 + 7]), ld(&W.A[64 + 21])), ld(&W.A[64 + 43])), ld(&W.A[64 + 57])));
                    const float pa = cabsf_(ps);
                    float2* __restrict__ bq = W.A + 128;              // [52] equalised SIG-B tones
                    float* __restrict__ coded = W.F + 192;            // [52]
                    for (int i = lane; i < 64; i += 32) {
                        const int d = L->binToDataNL[i];
                        if (d == 255) continue;
                        const cf q = cdivs(cmul(ld(&W.A[64 + i]), ps), pa);
                        bq[d] = st(q);
                        coded[L->deintNL[0][0][d]] = q.re;            // mapDeintVhtSigB20
                    }
                    __syncwarp();
                    uint8_t sb[26], enc[52];
                    unpack_bits(sig_viterbi_w(L, coded, 26, W, lane), sb, 26);
                    bcc_encode(sb, enc, 26);
                    double np = 0.0;
                    for (int i = 0; i < 52; i++) {                // procIntelVhtB20 + noise power :488-504
                        const cf q = ld(&bq[i]);
                        const cf e = mk(fsub(q.re, enc[L->deintNL[0][0][i]] ? 1.0f : -1.0f), q.im);
                        np += (double)fadd(fmul(e.re, e.re), fmul(e.im, e.im));
                    }
                    sssnr = (float)(log10(52.0 / np) * 10.0);
                    parse_vhtb(sb, &m);
                    const int nl = (llen * 8 + 22 + 23) / 24;
                    const bool ok = m.len >= 0 && m.len <= 4095 && m.nSS <= 2 && (nl * 80) >= (m.nSym * m.nSymSamp + 160 + 80 + m.nLTF * 80 + 80);
                    pos += need;
                    if (!ok) status = C8B_ST_FORMAT;
                    trellis = m.nSym * m.nDBPS;
                }
            } else {
                uint8_t hb[48];
                unpack_bits(sig_viterbi_w(L, llrht, 48, W, lane), hb, 48);
                if (check_ht(hb)) {                               // DEMOD_S_HT :180-205
                    parse_ht(hb, &m);
                    pos = 160;
                    const int need = 80 + m.nLTF * 80;
                    if (nsig - pos < need) status = C8B_ST_TRUNC;
                    else {
                        __syncwarp();
                        win(0, pos + 80);
                        fft64_groups(L, W.A, W, 1, lane);
                        for (int i = lane; i < 64; i += 32) HNL[i] = nl_null(i) ? make_float2(0.f, 0.f) : st(cdivs(ld(&W.A[i]), L->ltfNL[i]));
                        haveNL = true;
                        __syncwarp();
                        const int nl = (llen * 8 + 22 + 23) / 24;
                        const bool ok = m.len > 0 && m.len <= 4095 && m.nSS <= 2 && (nl * 80) >= (m.nSym * m.nSymSamp + 160 + 80 + m.nLTF * 80);
                        pos += need;
                        if (!ok) status = C8B_ST_FORMAT;
                        trellis = m.len * 8 + 22;
                    }
                } else legacy = true;
            }
        }
    }
    if (status == C8B_ST_OK && legacy) { parse_l(lmcs, llen, &m); trellis = m.len * 8 + 22; }
    if (status != C8B_ST_OK) { if (lane == 0) fr->status = status; return; }
    // DEMOD_S_WRTAG :221-277
    for (int i = lane; i < 64; i += 32) {
        const cf hh = (m.format == C8B_F_L) ? ld(&hl[i]) : (haveNL ? ld(&HNL[i]) : mk(0.f, 0.f));
        const bool used = (m.format == C8B_F_L) ? !l_null(i) : !nl_null(i);
        float2 r = make_float2(0.f, 0.f);
        if (used) { const double den = c8b::dadd(c8b::dmul((double)hh.re, (double)hh.re), c8b::dmul((double)hh.im, (double)hh.im)); r = make_float2((float)((double)hh.re / den), (float)(-(double)hh.im / den)); }
        hinv[i] = r;
    }
    int total = m.nSym * m.nCBPS;
    if (m.nSym == 0) { total = 1024; status = C8B_ST_NDP; }
    else if (m.nSS != 1) status = C8B_ST_FORMAT;                  // 2-stream frames need the demod2 path
    else if (pos + m.nSym * m.nSymSamp > nsig) status = C8B_ST_TRUNC;
    else if ((int64_t)total > llrStride) status = C8B_ST_OVERFLOW;
    if (status == C8B_ST_NDP && llrAll && llrStride >= 256) {
        // tag "mu2x1chan" (lib/demod_impl.cc:238-249): the 2 x 64 time samples of the two VHT-LTFs that the sounding branch
        // of nonLegacyChanEstimate keeps (:391-394); a one-stream NDP never fills them (zeros here)
        float2* __restrict__ o = reinterpret_cast<float2*>(llrAll + (size_t)sl * llrStride);
        for (int k = lane; k < 128; k += 32)
            o[k] = m.nSS != 1 ? st(rot.at(240 + C8B_SYM_SHIFT + (k & 63) + (k >> 6) * 80)) : make_float2(0.f, 0.f);
    }
    if (lane == 0) {
        fr->format = m.format; fr->mcs = m.mcs; fr->len = m.len; fr->cr = m.cr; fr->ampdu = m.ampdu;
        fr->nss = m.nSS; fr->nsym = m.nSym; fr->nsymsamp = m.nSymSamp; fr->ncbps = m.nCBPS; fr->ndbps = m.nDBPS;
        fr->trellis = trellis; fr->total = total; fr->data_off = pos;
        fr->sssnr0 = (m.format == C8B_F_VHT) ? sssnr : 0.f; fr->sssnr1 = 0.f;
        fr->status = status;
    }
}

// ---------------------------------------------------------------------------------------------------
// demod2 header states, two antennas (lib/demod2_impl.cc:72-277, :350-469, :632-758): warp version of demod_header2.
// Same control flow (format detection on antenna 0, HT / VHT parsers, SIG-B), every per-bin loop spread over the lanes
// (two bins per lane): the four LTF spectra, H from LTF1 +- LTF2, the VHT pilot-bin interpolation, the per-bin Gram
// matrix and its inverse (zero forcing as the reference, or MMSE with cfg.mmse), the folded weights
// (H^H H)^-1 H^H for k_demod2 in double, the two-stream SIG-B.  Sums whose order matters (pilot phase, SIG-B noise)
// are formed in the reference's order.
// ---------------------------------------------------------------------------------------------------
constexpr int FW2 = 2;                         // warps (= frames) per CTA: the workspace is 14 KB per warp
struct __align__(16) Ws2 {
    Ws w;                                       // spectra (w.A: 4 windows), soft bits (w.F), Viterbi / DFT scratch
    float2 H[256];                              // H[4 i + k], the reference's d_H_NL[i][k]
    float2 HI[256];                             // inverse Gram per bin
    float2 P[8];                                // pilot references pnl[4], pnl2[4]
};

__global__ void __launch_bounds__(FW2 * 32)
k_header2_w(const c8b_lut* __restrict__ lut, const float2* __restrict__ iq0, const float2* __restrict__ iq1, const int64_t* __restrict__ off,
            int nslots, int maxf, int mmse, c8b_frame* __restrict__ frames, const float2* __restrict__ chan, float2* __restrict__ hinvAll,
            float2* __restrict__ w2All, int64_t llrStride)
{
    __shared__ Ws2 ws[FW2];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int sl = blockIdx.x * FW2 + warp;
    if (sl >= nslots) return;
    Ws2& V = ws[warp];
    Ws& W = V.w;
    c8b_frame* __restrict__ fr = frames + sl;
    float2* __restrict__ hinv = hinvAll + (size_t)sl * 64;
    float2* __restrict__ w2 = w2All + (size_t)sl * 264;
    if (lane == 0) fr->llr_off = (int64_t)sl * llrStride;
    if (fr->status != C8B_ST_OK) return;
    const c8b_lut* L = lut;
    RotW r0, r1;
    r0.x = reinterpret_cast<const cf*>(iq0 + off[sl / maxf]) + fr->sync_idx + 224;
    r0.rad = fr->rad; r0.nsamp = fr->nsamp;
    r1 = r0; r1.x = reinterpret_cast<const cf*>(iq1 + off[sl / maxf]) + fr->sync_idx + 224;
    const int nsamp = fr->nsamp, lmcs = fr->l_mcs, llen = fr->l_len;
    const float sigma2 = mmse_sigma2(mmse, fr->snr, fr->rssi);
    const float2* __restrict__ hl = chan + (size_t)sl * 64;

    Mod m;
    m.format = m.sumu = m.ampdu = m.nSym = m.nSymSamp = m.nSD = m.nSP = m.nSS = m.nLTF = 0;
    m.mcs = m.len = m.mod = m.cr = m.nBPSCS = m.nDBPS = m.nCBPS = m.nCBPSS = 0;
    const int nsig = nsamp + 320;
    int pos = 0, trellis = 0, status = C8B_ST_OK;
    float sssnr0 = 0.f, sssnr1 = 0.f;
    bool legacy = lmcs > 0;
    for (int i = lane; i < 256; i += 32) { V.H[i] = make_float2(0.f, 0.f); V.HI[i] = make_float2(0.f, 0.f); }
    if (lane < 8) V.P[lane] = make_float2(0.f, 0.f);
    // window k (0..3) of 64 rotated samples of antenna a starting at stream index `start` -> W.A[64 k ..]
    auto win = [&](int k, int a, int start) {
        const RotW& r = a ? r1 : r0;
        for (int i = lane; i < 64; i += 32) W.A[64 * k + i] = st(r.at(start + C8B_SYM_SHIFT + i));
    };
    float* __restrict__ llrht = W.F, *llrvht = W.F + 96;           // 96 + 96 soft bits
    auto Hk = [&](int i, int k) { return ld(&V.H[4 * i + k]); };
    // zero-forcing / MMSE combine of one bin, the reference's operation order (:498-501)
    auto zf = [&](int i, cf a1, cf a2, cf& s1, cf& s2) {
        const cf t1 = cadd(cmul(a1, cconj(Hk(i, 0))), cmul(a2, cconj(Hk(i, 1))));
        const cf t2 = cadd(cmul(a1, cconj(Hk(i, 2))), cmul(a2, cconj(Hk(i, 3))));
        s1 = cadd(cmul(t1, ld(&V.HI[4 * i + 0])), cmul(t2, ld(&V.HI[4 * i + 2])));
        s2 = cadd(cmul(t1, ld(&V.HI[4 * i + 1])), cmul(t2, ld(&V.HI[4 * i + 3])));
    };
    auto chan_estimate = [&](int start) {                          // nonLegacyChanEstimate :350-469
        __syncwarp();
        if (m.nSS == 1) {
            if (m.nLTF == 1) {
                win(0, 0, start);
                fft64_groups(L, W.A, W, 1, lane);
                for (int i = lane; i < 64; i += 32) if (!nl_null(i)) V.H[4 * i] = st(cdivs(ld(&W.A[i]), L->ltfNL[i]));
            }
        } else if (m.nSS == 2) {
            win(0, 0, start); win(1, 1, start); win(2, 0, start + 80); win(3, 1, start + 80);
            fft64_groups(L, W.A, W, 4, lane);
            for (int i = lane; i < 64; i += 32) {
                if (nl_null(i)) continue;
                const float l2 = fmul(L->ltfNL[i], 0.5f);         // LTF_NL_28_F_FLOAT2
                const cf f1 = ld(&W.A[i]), f2 = ld(&W.A[64 + i]), f12 = ld(&W.A[128 + i]), f22 = ld(&W.A[192 + i]);
                V.H[4 * i + 0] = st(cscale(csub(f1, f12), l2)); V.H[4 * i + 1] = st(cscale(csub(f2, f22), l2));
                V.H[4 * i + 2] = st(cscale(cadd(f1, f12), l2)); V.H[4 * i + 3] = st(cscale(cadd(f2, f22), l2));
            }
            __syncwarp();
            const int pb[4] = { 7, 21, 43, 57 }, slot[4] = { 2, 3, 0, 1 };
            if (m.format == C8B_F_VHT && lane < 16) {             // pilot tones interpolated :391-409
                const int q = lane >> 2, k = lane & 3;
                V.H[4 * pb[q] + k] = st(cdivs(cadd(Hk(pb[q] - 1, k), Hk(pb[q] + 1, k)), 2.0f));
            }
            __syncwarp();
            for (int i = lane; i < 64; i += 32) {
                if (nl_null(i)) continue;
                const cf h0 = Hk(i, 0), h1 = Hk(i, 1), h2 = Hk(i, 2), h3 = Hk(i, 3);
                const cf a = cadd(cmul(h0, cconj(h0)), cmul(h1, cconj(h1)));
                const cf b = cadd(cmul(h0, cconj(h2)), cmul(h1, cconj(h3)));
                const cf c = cadd(cmul(h2, cconj(h0)), cmul(h3, cconj(h1)));
                const cf d = cadd(cmul(h2, cconj(h2)), cmul(h3, cconj(h3)));
                cf hi[4];
                gram_inverse(a, b, c, d, sigma2, hi);
#pragma unroll
                for (int k = 0; k < 4; k++) V.HI[4 * i + k] = st(hi[k]);
            }
            __syncwarp();
            if (lane < 4) {
                cf t1, t2;
                zf(pb[lane], ld(&W.A[pb[lane]]), ld(&W.A[64 + pb[lane]]), t1, t2);
                if (lane == 3) { t1 = mk(-t1.re, -t1.im); t2 = mk(-t2.re, -t2.im); }
                V.P[slot[lane]] = st(cconj(t1)); V.P[4 + slot[lane]] = st(cconj(t2));
            }
        }
        __syncwarp();
    };

    if (!legacy) {                                                // DEMOD_S_FORMAT
        if (nsig < 160) status = C8B_ST_TRUNC;
        else {
            __syncwarp();
            win(0, 0, 0); win(1, 0, 80);
            fft64_groups(L, W.A, W, 2, lane);
            {   // procNLSigDemodDeint (lib/cloud80211phy.cc:629-648) on antenna 0
                cf a1 = cdiv(ld(&W.A[7]), ld(&hl[7])), a2 = cdiv(ld(&W.A[64 + 7]), ld(&hl[7]));
                a1 = csub(a1, cdiv(ld(&W.A[21]), ld(&hl[21]))); a2 = csub(a2, cdiv(ld(&W.A[64 + 21]), ld(&hl[21])));
                a1 = cadd(a1, cdiv(ld(&W.A[43]), ld(&hl[43]))); a2 = cadd(a2, cdiv(ld(&W.A[64 + 43]), ld(&hl[43])));
                a1 = cadd(a1, cdiv(ld(&W.A[57]), ld(&hl[57]))); a2 = cadd(a2, cdiv(ld(&W.A[64 + 57]), ld(&hl[57])));
                const cf p1 = cconj(a1), p2 = cconj(a2);
                const float m1 = cabsf_(p1), m2 = cabsf_(p2);
                for (int i = lane; i < 64; i += 32) {
                    const int d = L->sigDemap[i];
                    if (d < 0) continue;
                    const cf h = ld(&hl[i]);
                    const cf q1 = cdivs(cmul(cdiv(ld(&W.A[i]), h), p1), m1);
                    const cf q2 = cdivs(cmul(cdiv(ld(&W.A[64 + i]), h), p2), m2);
                    llrht[d] = q1.im; llrht[d + 48] = q2.im;
                    llrvht[d] = q1.re; llrvht[d + 48] = q2.im;
                }
            }
            __syncwarp();
            uint8_t vb[48];
            unpack_bits(sig_viterbi_w(L, llrvht, 48, W, lane), vb, 48);
            if (check_vhta(vb)) {                                 // DEMOD_S_VHT
                parse_vhta(vb, &m);
                pos = 160;
                const int need = 80 + m.nLTF * 80 + 80;
                if (nsig - pos < need) status = C8B_ST_TRUNC;
                else {
                    chan_estimate(pos + 80);
                    // vhtSigBDemod :632-758
                    const int stb = pos + 80 + m.nLTF * 80;
                    float2* __restrict__ q0 = W.B, *q1 = W.B + 64;    // [52] equalised SIG-B tones per stream
                    float* __restrict__ coded = W.F + 192;            // [52]
                    uint8_t sb[26];
                    bool have = true;
                    if (m.nSS == 1) {
                        win(0, 0, stb);
                        fft64_groups(L, W.A, W, 1, lane);
                        for (int i = lane; i < 64; i += 32) if (!nl_null(i)) W.A[64 + i] = st(cdiv(ld(&W.A[i]), Hk(i, 0)));
                        __syncwarp();
                        const cf ps = cconj(cadd(cadd(csub(ld(&W.A[64 + 7]), ld(&W.A[64 + 21])), ld(&W.A[64 + 43])), ld(&W.A[64 + 57])));
                        const float pa = cabsf_(ps);
                        for (int i = lane; i < 64; i += 32) {
                            const int d = L->binToDataNL[i];
                            if (d == 255) continue;
                            const cf q = cdivs(cmul(ld(&W.A[64 + i]), ps), pa);
                            q0[d] = st(q);
                            coded[L->deintNL[0][0][d]] = q.re;
                        }
                    } else if (m.nSS == 2) {
                        win(0, 0, stb); win(1, 1, stb);
                        fft64_groups(L, W.A, W, 2, lane);
                        for (int i = lane; i < 64; i += 32) {
                            if (nl_null(i)) continue;
                            cf s1, s2;
                            zf(i, ld(&W.A[i]), ld(&W.A[64 + i]), s1, s2);
                            W.A[128 + i] = st(s1); W.A[192 + i] = st(s2);
                        }
                        __syncwarp();
                        const float2* s1 = W.A + 128, *s2 = W.A + 192;
                        cf acc = cmul(ld(&s1[7]), ld(&V.P[2]));
                        acc = csub(acc, cmul(ld(&s1[21]), ld(&V.P[3]))); acc = cadd(acc, cmul(ld(&s1[43]), ld(&V.P[0]))); acc = cadd(acc, cmul(ld(&s1[57]), ld(&V.P[1])));
                        acc = cadd(acc, cmul(ld(&s2[7]), ld(&V.P[6]))); acc = csub(acc, cmul(ld(&s2[21]), ld(&V.P[7])));
                        acc = cadd(acc, cmul(ld(&s2[43]), ld(&V.P[4]))); acc = cadd(acc, cmul(ld(&s2[57]), ld(&V.P[5])));
                        const cf ps = cconj(acc);
                        const float pa = cabsf_(ps);
                        for (int i = lane; i < 64; i += 32) {
                            const int d = L->binToDataNL[i];
                            if (d == 255) continue;
                            const cf a = cdivs(cmul(ld(&s1[i]), ps), pa), b = cdivs(cmul(ld(&s2[i]), ps), pa);
                            q0[d] = st(a); q1[d] = st(b);
                            coded[L->deintNL[0][0][d]] = fdiv(fadd(a.re, b.re), 2.0f);
                        }
                    } else have = false;
                    __syncwarp();
                    if (have) {
                        uint8_t enc[52];
                        unpack_bits(sig_viterbi_w(L, coded, 26, W, lane), sb, 26);
                        bcc_encode(sb, enc, 26);
                        double n0 = 0.0, n1 = 0.0;
                        for (int i = 0; i < 52; i++) {
                            const float ref = enc[L->deintNL[0][0][i]] ? 1.0f : -1.0f;
                            const cf a = ld(&q0[i]);
                            const cf e0 = mk(fsub(a.re, ref), a.im);
                            n0 += (double)fadd(fmul(e0.re, e0.re), fmul(e0.im, e0.im));
                            if (m.nSS == 2) { const cf b = ld(&q1[i]); const cf e1 = mk(fsub(b.re, ref), b.im); n1 += (double)fadd(fmul(e1.re, e1.re), fmul(e1.im, e1.im)); }
                        }
                        sssnr0 = (float)(log10(52.0 / n0) * 10.0);
                        if (m.nSS == 2) sssnr1 = (float)(log10(52.0 / n1) * 10.0);
                    } else {
                        for (int i = 0; i < 26; i++) sb[i] = 0;
                    }
                    parse_vhtb(sb, &m);
                    const int nl = (llen * 8 + 22 + 23) / 24;
                    const bool ok = m.len > 0 && m.len <= 4095 && m.nSS <= 2 && (nl * 80) >= (m.nSym * m.nSymSamp + 160 + 80 + m.nLTF * 80 + 80);
                    pos += need;
                    if (!ok) status = C8B_ST_FORMAT;
                    trellis = m.nSym * m.nDBPS;
                }
            } else {
                uint8_t hb[48];
                unpack_bits(sig_viterbi_w(L, llrht, 48, W, lane), hb, 48);
                if (check_ht(hb)) {                               // DEMOD_S_HT
                    parse_ht(hb, &m);
                    pos = 160;
                    const int need = 80 + m.nLTF * 80;
                    if (nsig - pos < need) status = C8B_ST_TRUNC;
                    else {
                        chan_estimate(pos + 80);
                        const int nl = (llen * 8 + 22 + 23) / 24;
                        const bool ok = m.len > 0 && m.len <= 4095 && m.nSS <= 2 && (nl * 80) >= (m.nSym * m.nSymSamp + 160 + 80 + m.nLTF * 80);
                        pos += need;
                        if (!ok) status = C8B_ST_FORMAT;
                        trellis = m.len * 8 + 22;
                    }
                } else legacy = true;
            }
        }
    }
    if (status == C8B_ST_OK && legacy) { parse_l(lmcs, llen, &m); trellis = m.len * 8 + 22; }
    if (status != C8B_ST_OK) { if (lane == 0) fr->status = status; return; }
    __syncwarp();
    for (int i = lane; i < 64; i += 32) {
        const cf hh = (m.format == C8B_F_L) ? ld(&hl[i]) : Hk(i, 0);
        const bool used = (m.format == C8B_F_L) ? !l_null(i) : !nl_null(i);
        const double den = c8b::dadd(c8b::dmul((double)hh.re, (double)hh.re), c8b::dmul((double)hh.im, (double)hh.im));
        hinv[i] = (used && m.nSS == 1) ? make_float2((float)((double)hh.re / den), (float)(-(double)hh.im / den)) : make_float2(0.f, 0.f);
        float2 wv[4] = { make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f) };
        if (m.nSS == 2 && !nl_null(i)) {
            const cd h0 = dconj(tod(Hk(i, 0))), h1 = dconj(tod(Hk(i, 1))), h2 = dconj(tod(Hk(i, 2))), h3 = dconj(tod(Hk(i, 3)));
            const cd i0 = tod(ld(&V.HI[4 * i + 0])), i1 = tod(ld(&V.HI[4 * i + 1])), i2 = tod(ld(&V.HI[4 * i + 2])), i3 = tod(ld(&V.HI[4 * i + 3]));
            wv[0] = st(tof(dcadd(dcmul(h0, i0), dcmul(h2, i2))));    // stream 0 <- antenna 0
            wv[1] = st(tof(dcadd(dcmul(h1, i0), dcmul(h3, i2))));    // stream 0 <- antenna 1
            wv[2] = st(tof(dcadd(dcmul(h0, i1), dcmul(h2, i3))));    // stream 1 <- antenna 0
            wv[3] = st(tof(dcadd(dcmul(h1, i1), dcmul(h3, i3))));    // stream 1 <- antenna 1
        }
#pragma unroll
        for (int k = 0; k < 4; k++) w2[4 * i + k] = wv[k];
    }
    if (lane < 8) w2[256 + lane] = V.P[lane];
    if (m.nSS != 1 && m.nSS != 2) status = C8B_ST_FORMAT;
    else if (m.nSS == 1 && m.format != C8B_F_L && m.nLTF != 1) status = C8B_ST_FORMAT;   // reference leaves H unset (:355-372)
    else if (pos + m.nSym * m.nSymSamp > nsig) status = C8B_ST_TRUNC;
    const int total = m.nSym * m.nCBPS;
    if (status == C8B_ST_OK && (int64_t)total > llrStride) status = C8B_ST_OVERFLOW;
    if (lane == 0) {
        fr->format = m.format; fr->mcs = m.mcs; fr->len = m.len; fr->cr = m.cr; fr->ampdu = m.ampdu;
        fr->nss = m.nSS; fr->nsym = m.nSym; fr->nsymsamp = m.nSymSamp; fr->ncbps = m.nCBPS; fr->ndbps = m.nDBPS;
        fr->trellis = trellis; fr->total = total; fr->data_off = pos;
        fr->sssnr0 = (m.format == C8B_F_VHT) ? sssnr0 : 0.f;
        fr->sssnr1 = (m.format == C8B_F_VHT && m.nSS == 2) ? sssnr1 : 0.f;
        fr->status = status;
    }
}

}  // namespace

void c8b_launch_header2_w(const c8b_lut* lut, const float2* iq0, const float2* iq1, const int64_t* d_off, int nitems, int maxf, int mmse,
                          c8b_frame* frames, const float2* chan, float2* hinv, float2* w2, int64_t llrStride, cudaStream_t st)
{
    if (nitems <= 0) return;
    const int ns = nitems * maxf;
    k_header2_w<<<(ns + FW2 - 1) / FW2, FW2 * 32, 0, st>>>(lut, iq0, iq1, d_off, ns, maxf, mmse, frames, chan, hinv, w2, llrStride);
}

// ---------------------------------------------------------------------------------------------------
// One event of one block per launch: the kernels behind the per-block entry points (csrc/blocks.cu, c8b_blk_work), the
// same warp routines as the batch path.  Result layouts are those of c8b_blocks::SyncRes / SignalRes (blocks.h); the result
// pointers may be mapped host memory (a 20-byte / 1 KB posted write instead of a device-to-host copy).
// ---------------------------------------------------------------------------------------------------
namespace {
__global__ void __launch_bounds__(32)
k_one_sync_w(const float2* __restrict__ sig, float2 conj, int32_t* __restrict__ res)
{
    __shared__ Ws w;
    const SyncOut o = sync_at_w(reinterpret_cast<const cf*>(sig), mk(conj.x, conj.y), w, threadIdx.x);
    if (threadIdx.x == 0) {
        res[0] = o.ok; res[1] = o.mIndex;
        reinterpret_cast<float*>(res)[2] = o.rad; reinterpret_cast<float*>(res)[3] = o.snr; reinterpret_cast<float*>(res)[4] = o.rssi;
    }
}

// L-SIG of one sync flag (in = >= 224 samples from the flag) and, in the same launch, the S_COPY loop of the call that
// found it (lib/signal_impl.cc:108-192): ncopy samples from in + 224 rotated by the CFO phase into out0 (out1: antenna 1
// of signal2 from in1 + 224).  The host decides afterwards how many of them the block's accounting hands on (it knows
// nsamp only from this kernel's result); ncopy = 0: the L-SIG alone.
__global__ void __launch_bounds__(128)
k_one_signal_w(const c8b_lut* __restrict__ lut, const float2* __restrict__ in, const float2* __restrict__ in1, float rad, int32_t* __restrict__ res,
               float2* __restrict__ out0, float2* __restrict__ out1, int ncopy)
{
    __shared__ Ws w;
    if (threadIdx.x < 32) {
        int mcs = 0, len = 0, nsamp = 0;
        const int ok = signal_at_w(lut, reinterpret_cast<const cf*>(in), rad, reinterpret_cast<float2*>(res + 4), &mcs, &len, &nsamp, w, threadIdx.x);
        if (threadIdx.x == 0) { res[0] = ok; res[1] = mcs; res[2] = len; res[3] = nsamp; }
    } else {
        for (int i = threadIdx.x - 32; i < ncopy; i += 96) {
            const cf ph = cis(fmul((float)(i + 224), rad));          // lib/signal_impl.cc:172-173, d_nSampleCopied = i
            out0[i] = st(cmul(ld(&in[224 + i]), ph));
            if (in1) out1[i] = st(cmul(ld(&in1[224 + i]), ph));
        }
    }
}

// The trigger FSM (lib/trigger_impl.cc:59-117) over one call's samples, a bitmap word (32 samples, lane = sample) at a
// time, continued from the state the previous call left: the FSM changes course only where a plateau reaches its 21st
// sample (count-down armed) and where the count-down ends (0x01); between those the word is applied in closed form, and
// the 0x02 flags ("new maximum of the plateau") are the lanes that beat the running prefix maximum of their run.
__global__ void __launch_bounds__(32)
k_one_trigger_w(TrigState* __restrict__ state, const float* __restrict__ in, int n, uint8_t* __restrict__ out)
{
    const int lane = threadIdx.x;
    TrigState s = *state;
    constexpr int WB = 8;                                         // words fetched together: the loads of a block are independent,
    for (int blk = 0; blk < n; blk += 32 * WB) {                   // only the FSM that walks them is serial
        float pvs[WB];
        uint32_t aboves[WB];
#pragma unroll
        for (int w = 0; w < WB; w++) {
            const int i = blk + 32 * w + lane;
            pvs[w] = i < n ? in[i] : 0.f;
        }
#pragma unroll
        for (int w = 0; w < WB; w++) aboves[w] = __ballot_sync(FULL, blk + 32 * w + lane < n && pvs[w] > 0.3f);
#pragma unroll
        for (int w = 0; w < WB; w++) {
            const int base = blk + 32 * w;
            if (base >= n) break;
            const int i = base + lane, kmax = min(32, n - base);
            const float pv = pvs[w];
            const uint32_t above = aboves[w];
            uint32_t fl = 0;
            if (above == 0u && kmax == 32 && !s.fPlateau) {         // a quiet word with no count-down running: the reset state
                s.nPlateau = 0; s.fPlateauEnd = 0; s.conjAc = 0.0f;
                out[i] = 0;
                continue;
            }
            for (int k = 0; k < kmax;) {
                const uint32_t rest = above >> k;
                if (!(rest & 1u)) {                                   // samples k .. k + gap - 1 below the threshold (:96-101)
                    const int gap = rest ? __ffs(rest) - 1 : kmax - k;
                    s.nPlateau = 0; s.fPlateauEnd = 0; s.conjAc = 0.0f;
                    if (s.fPlateau) {
                        if (s.countDown <= gap) { if (lane == k + s.countDown - 1) fl |= 1u; s.countDown = 0; s.fPlateau = 0; }
                        else s.countDown -= gap;
                    }
                    k += gap;
                    continue;
                }
                const uint32_t inv = ~rest;
                const int P = min(inv ? __ffs(inv) - 1 : 32, kmax - k);      // run above the threshold: samples k .. k + P - 1
                const bool inRun = lane >= k && lane < k + P;
                float incl = inRun ? pv : -1.0f;                          // inclusive prefix maximum over the run (preac >= 0)
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { const float t = __shfl_up_sync(FULL, incl, o); if (lane >= o) incl = fmaxf(incl, t); }
                float excl = __shfl_up_sync(FULL, incl, 1);
                if (lane == 0) excl = -1.0f;
                if (inRun && pv > fmaxf(s.conjAc, excl)) fl |= 2u;         // :82-87
                const float runMax = __shfl_sync(FULL, incl, k + P - 1);
                int done = 0;                                             // samples of the run already accounted for
                if (s.fPlateau) {                                         // a count-down from an earlier plateau runs through (:103-111)
                    if (s.countDown <= P) { if (lane == k + s.countDown - 1) fl |= 1u; done = s.countDown; s.countDown = 0; s.fPlateau = 0; }
                    else { s.countDown -= P; done = P; }
                }
                if (!s.fPlateau && s.fPlateauEnd == 0 && done < P) {      // armed at the first sample with nPlateau > 20 (:88-93)
                    const int r = max(done + 1, 21 - s.nPlateau);
                    if (r <= P) { s.fPlateau = 1; s.fPlateauEnd = 1; s.countDown = 79 - (P - r); }
                }
                s.nPlateau += P;
                s.conjAc = fmaxf(s.conjAc, runMax);
                k += P;
            }
            if (i < n) out[i] = (uint8_t)fl;
        }
    }
    __syncwarp();
    if (lane == 0) *state = s;
}
}  // namespace

void c8b_launch_one_sync(const float2* d_sig, float conj_re, float conj_im, void* res, cudaStream_t st)
{
    k_one_sync_w<<<1, 32, 0, st>>>(d_sig, make_float2(conj_re, conj_im), reinterpret_cast<int32_t*>(res));
}
void c8b_launch_one_signal(const c8b_lut* lut, const float2* d_in, const float2* d_in1, float rad, void* res, float2* d_out0, float2* d_out1,
                           int ncopy, cudaStream_t st)
{
    k_one_signal_w<<<1, 128, 0, st>>>(lut, d_in, d_in1, rad, reinterpret_cast<int32_t*>(res), d_out0, d_out1, ncopy);
}
void c8b_launch_one_trigger(void* d_state, const float* d_in, int n, uint8_t* d_out, cudaStream_t st)
{
    k_one_trigger_w<<<1, 32, 0, st>>>(reinterpret_cast<TrigState*>(d_state), d_in, n, d_out);
}

void c8b_launch_detect_w(const c8b_lut* lut, const float2* iq, const int64_t* d_off, const int32_t* d_len, int nitems, int itemBase,
                         int maxf, int64_t outBase, const float* preac, const uint32_t* mask, int maskStride, c8b_frame* frames,
                         float2* chan, c8b_scan* scans, cudaStream_t st)
{
    if (nitems <= 0) return;
    // one warp per item, frames of an item found serially: the batch path (many items).  Stream windows (scans) and
    // long captures go through c8b_launch_detect_multi.
    (void)scans;
    k_detect_w<<<(nitems + FW - 1) / FW, FW * 32, 0, st>>>(lut, iq, d_off, d_len, nitems, itemBase, maxf, outBase, preac, mask, maskStride,
                                                            frames, chan);
}

// staged entry point for the trigger scan alone (c8b_trigger_events): bitmap of a given preac array, then k_trig_scan
namespace {
__global__ void k_preac_mask(const float* __restrict__ preac, int n, uint32_t* __restrict__ mask)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t m = __ballot_sync(FULL, i < n && preac[i < n ? i : 0] > 0.3f);
    if ((threadIdx.x & 31) == 0 && i < n) mask[i >> 5] = m;
}
__global__ void k_cand_export(const Cand* __restrict__ cands, const CandHead* __restrict__ heads, const uint8_t* __restrict__ tabs, int nsegMax, int cap,
                              int32_t* __restrict__ out)
{
    // one item: the segments' candidate lists concatenated in stream order (single thread: a test entry point)
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const int nseg = reinterpret_cast<const SegTab*>(tabs)->nseg;
    int n = 0, overflow = 0, nTrig = 0, safeEnd = 0;
    bool stop = false;
    for (int sg = 0; sg < nseg && !stop; sg++) {
        const CandHead hd = heads[sg];
        nTrig += hd.nTrig;
        safeEnd = hd.safeEnd;
        for (int c = 0; c < hd.n; c++, n++) {
            const Cand& q = cands[(size_t)sg * C8B_SEG_CAP + c];
            if (n < cap) { out[4 + 4 * n + 0] = q.trig; out[4 + 4 * n + 1] = q.latch; out[4 + 4 * n + 2] = q.safe; out[4 + 4 * n + 3] = q.stall; }
            else overflow = 1;
            if (q.stall) { safeEnd = q.safe; stop = true; n++; break; }   // the serial scan ends at the first stalled trigger
        }
        if (hd.overflow) { overflow = 1; stop = true; }
    }
    out[0] = n < cap ? n : cap; out[1] = safeEnd; out[2] = overflow; out[3] = nTrig;
}
}  // namespace

static int seg_words_cfg()
{
    // C8B_SEG_WORDS (environment): nominal segment length in bitmap words, for tests that want many cuts in a short capture
    static int v = -1;
    if (v < 0) { const char* e = getenv("C8B_SEG_WORDS"); v = e ? atoi(e) : 0; if (v < 8) v = C8B_SEG_WORDS; }
    return v;
}
static int nseg_max_for(int maxLen) { return maxLen / (32 * seg_words_cfg()) + 2; }

size_t c8b_detect_multi_scratch(int nitems, int maxLen)
{
    const int nsm = nseg_max_for(maxLen);
    return (size_t)nitems * (segtab_bytes(nsm) + (size_t)nsm * (sizeof(CandHead) + (size_t)C8B_SEG_CAP * sizeof(Cand))) + 512;
}

namespace {
struct MultiScratch { uint8_t* tabs; CandHead* heads; Cand* cands; int nsegMax; };
MultiScratch carve(void* scratch, int nitems, int maxLen)
{
    MultiScratch m;
    m.nsegMax = nseg_max_for(maxLen);
    uint8_t* p = reinterpret_cast<uint8_t*>(scratch);
    m.tabs = p;
    p += ((size_t)nitems * segtab_bytes(m.nsegMax) + 255) & ~(size_t)255;
    m.heads = reinterpret_cast<CandHead*>(p);
    p += ((size_t)nitems * m.nsegMax * sizeof(CandHead) + 255) & ~(size_t)255;
    m.cands = reinterpret_cast<Cand*>(p);
    return m;
}
}  // namespace

// d_item: {int64 off = 0; int32 len = n} already on the device; out = 4 header ints + 4 ints per event
void c8b_launch_trigger_events(const float* d_preac, int n, const int64_t* d_off, const int32_t* d_len, uint32_t* d_mask, const c8b_scan* d_scan,
                               void* scratch, int cap, int32_t* d_out, cudaStream_t st)
{
    const MultiScratch m = carve(scratch, 1, n);
    const int maskStride = (n + 31) / 32 + 1;
    k_preac_mask<<<(n + 255) / 256, 256, 0, st>>>(d_preac, n, d_mask);
    k_seg_cuts<<<1, 256, 0, st>>>(d_len, 1, d_mask, maskStride, d_scan, m.tabs, m.nsegMax, seg_words_cfg());
    k_trig_scan<<<m.nsegMax, 32, 0, st>>>(d_len, 1, 0, d_off, d_preac, d_mask, maskStride, m.tabs, m.nsegMax, m.cands, m.heads);
    k_cand_export<<<1, 32, 0, st>>>(m.cands, m.heads, m.tabs, m.nsegMax, cap, d_out);
}

// few long items with many frames each: segment cuts -> trigger scans in parallel -> per-trigger sync / signal in parallel -> accept rules
void c8b_launch_detect_multi(const c8b_lut* lut, const float2* iq, const int64_t* d_off, const int32_t* d_len, int nitems, int itemBase,
                             int maxf, int64_t outBase, const float* preac, const uint32_t* mask, int maskStride, c8b_frame* frames,
                             float2* chan, c8b_scan* scans, void* scratch, int maxLen, cudaStream_t st)
{
    if (nitems <= 0) return;
    const MultiScratch m = carve(scratch, nitems, maxLen);
    k_seg_cuts<<<nitems, 256, 0, st>>>(d_len, nitems, mask, maskStride, scans, m.tabs, m.nsegMax, seg_words_cfg());
    k_trig_scan<<<nitems * m.nsegMax, 32, 0, st>>>(d_len, nitems, outBase, d_off, preac, mask, maskStride, m.tabs, m.nsegMax, m.cands, m.heads);
    const int64_t warps = (int64_t)nitems * m.nsegMax * C8B_SEG_CAP;
    k_cand_eval<<<(unsigned)((warps + FW - 1) / FW), FW * 32, 0, st>>>(lut, iq, d_off, d_len, nitems, m.cands, m.heads, m.tabs, m.nsegMax);
    k_cand_accept<<<nitems, 32, 0, st>>>(d_len, nitems, itemBase, maxf, m.cands, m.heads, m.tabs, m.nsegMax, frames, chan, scans);
}

void c8b_launch_header_w(const c8b_lut* lut, const float2* iq, const int64_t* d_off, int nitems, int maxf, int mupos, c8b_frame* frames,
                         const float2* chan, float2* hinv, int64_t llrStride, float* llr, cudaStream_t st)
{
    if (nitems <= 0) return;
    const int ns = nitems * maxf;
    k_header_w<<<(ns + FW - 1) / FW, FW * 32, 0, st>>>(lut, iq, d_off, ns, maxf, mupos, frames, chan, hinv, llrStride, llr);
}
