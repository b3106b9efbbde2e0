// k_tx.cu -- on-device synthesiser of 20 MHz one- and two-stream 802.11a/g/n/ac transmit waveforms: the reference's
// encode -> modulation -> IFFT/CP -> pad chain (lib/encode_impl.cc:130-241, lib/modulation_impl.cc, lib/pad_impl.cc:37-80,
// lib/cloud80211phy.cc:2594-3161) in the form of its Python twin tools/phy80211.py (genFromMpdu / genFromAmpdu /
// genFinalSig), which is the generator every receive test of the reference and of this repo is fed from -- the waveforms
// must equal that generator's sample for sample (tests/test_gpu_tx.py against tests/golden/frames_siso.npz).
//
// Nothing is sequential: a frame is a row of 80-sample slots, every slot is the inverse DFT of a 64-bin spectrum, and
// every spectrum bin is a pure function of the frame descriptor and the PSDU bytes --
//   data bit x          = PSDU / service / pad bit XOR the period-127 scrambler sequence (tail bits forced to 0),
//   coded bit (t, o)    = parity of data bits t-6..t under 0155 / 0117 (tools/phy80211header.py:763-798),
//   punctured stream u  -> mother-code index in closed form, interleaved position j -> coded index through the receive
//                          LUT (the deinterleave map is the interleaver read backwards),
// so one CTA of 64 threads builds the 64 bins of a slot, transforms them and writes the 80 samples.  The windowing of
// procConcat2Symbol (first sample of a field and last sample of the field before it halved,
// tools/phy80211header.py:894-901), tone scaling (:966-967), pilots (polarity sequence, per-symbol rotation for HT / VHT,
// tools/phy80211.py:770-812) and the CFO rotation of genFinalSig (:832-838) are folded into the same pass.
// Two spatial streams (HT MCS 8-15, VHT with two space-time streams; lib/encode2_impl.cc, lib/modulation2_impl.cc,
// tools/phy80211.py:223-235,459-510,712-760): stream k goes to antenna k (direct mapping).  What changes per stream is
// closed-form too: the stream parser (blocks of s = max(nBPSCS / 2, 1) coded bits alternate between the streams), the
// second stream's interleaver rotation (the receive LUT deintNL[1]), its pilot pattern (HT), the LTF signs of the P matrix
// (R on the VHT pilot tones), the 1 / sqrt(2) power split and the cyclic shifts (200 ns = 4 samples on the legacy part,
// 400 ns = 8 samples from the HT / VHT-STF on) -- a cyclic shift is an index offset into the inverse DFT's output.
#include "common.cuh"

namespace {

constexpr int TXS = 4;              // slots per CTA (64 threads each)

struct TxPlan {                     // per frame, filled by k_tx_plan
    int32_t format, mcs, nbpsc, cr, ncbps, ndbps, nsym, nslots;
    int32_t nss, npre;              // spatial streams (1, 2); slots in front of the DATA field
    int32_t psdu_len;               // bytes handed in (MPDU, or A-MPDU for VHT)
    int32_t nbits;                  // nsym * ndbps
    int32_t tail0;                  // L / HT: first tail bit (16 + 8 * psdu_len)
    int32_t npadeof;                // VHT: EOF padding delimiters
    uint32_t lsig, sigb, svc_crc;   // L-SIG 24 bits, VHT-SIG-B 26 bits, SIG-B CRC carried in the service field
    uint64_t sig48;                 // HT-SIG / VHT-SIG-A 48 bits
    int64_t psdu_off, out_off;
    int64_t q_off;                  // MU-MIMO: first element of the frame's 64 x 2 x 2 spatial mapping matrices
    double cfo_step;                // rad / sample
};

// mcs of a descriptor: legacy 0-7; HT 0-15 (8-15: two streams); VHT: MCS 0-8 + 16 for two space-time streams
__host__ __device__ inline int tx_nss(int format, int mcs)
{
    if (format == C8B_F_HT) return mcs >= 8 ? 2 : 1;
    if (format == C8B_F_VHT) return mcs >= 16 ? 2 : 1;
    return 1;
}
__host__ __device__ inline int tx_rate(int format, int mcs, int* nbpsc, int* cr)
{
    if (format == C8B_F_HT && mcs >= 8 && mcs <= 15) mcs -= 8;
    if (format == C8B_F_VHT && mcs >= 16 && mcs <= 24) mcs -= 16;
    // legacy: signalParserL table; HT 0-7 / VHT 0-8: modulation class of tools/phy80211header.py:258-365
    const int lb[8] = { 1, 1, 2, 2, 4, 4, 6, 6 };
    const int lc[8] = { C8B_CR_12, C8B_CR_34, C8B_CR_12, C8B_CR_34, C8B_CR_12, C8B_CR_34, C8B_CR_23, C8B_CR_34 };
    const int nb[10] = { 1, 2, 2, 4, 4, 6, 6, 6, 8, 8 };
    const int nc[10] = { C8B_CR_12, C8B_CR_12, C8B_CR_34, C8B_CR_12, C8B_CR_34, C8B_CR_23, C8B_CR_34, C8B_CR_56, C8B_CR_34, C8B_CR_56 };
    if (format == C8B_F_L) { if (mcs < 0 || mcs > 7) return 0; *nbpsc = lb[mcs]; *cr = lc[mcs]; return 1; }
    if (format == C8B_F_HT) { if (mcs < 0 || mcs > 7) return 0; *nbpsc = nb[mcs]; *cr = nc[mcs]; return 1; }
    if (format == C8B_F_VHT) { if (mcs < 0 || mcs > 8) return 0; *nbpsc = nb[mcs]; *cr = nc[mcs]; return 1; }
    return 0;
}
__host__ __device__ inline int tx_ndbps(int ncbps, int cr)
{
    return cr == C8B_CR_12 ? ncbps / 2 : cr == C8B_CR_23 ? ncbps * 2 / 3 : cr == C8B_CR_34 ? ncbps * 3 / 4 : ncbps * 5 / 6;
}
// number of OFDM symbols and of 80-sample slots (tools/phy80211header.py:436-512)
__host__ __device__ inline int tx_geometry(int format, int mcs, int len, int* nsym, int* nslots)
{
    int nbpsc = 0, cr = 0;
    if (!tx_rate(format, mcs, &nbpsc, &cr) || len < 0 || len > 4095 || (len == 0 && format != C8B_F_VHT)) return 0;
    const int nss = tx_nss(format, mcs);
    const int ndbps = tx_ndbps((format == C8B_F_L ? 48 : 52) * nbpsc * nss, cr);
    const int bits = 8 * len + 16 + 6;
    *nsym = len == 0 ? 0 : (bits + ndbps - 1) / ndbps;
    *nslots = (format == C8B_F_L ? 5 : format == C8B_F_HT ? 8 + nss : 9 + nss) + *nsym;      // nLTF = nss
    return 1;
}

__host__ __device__ inline uint32_t crc8_bits(uint64_t bits, int n)           // genBitBitCrc8, tools/phy80211header.py:87-100
{
    uint32_t c = 0xff;                                                      // c[i] = bit i
    for (int k = 0; k < n; k++) {
        const uint32_t b = (uint32_t)(bits >> k) & 1u, f = b ^ ((c >> 7) & 1u);
        c = ((c << 1) & 0xff) ^ (f ? 0x07u : 0u);                           // next_c[0] = f, [1] = f ^ c0, [2] = f ^ c1, [i] = c[i-1]
    }
    uint32_t out = 0;                                                       // [1 - b for b in c[::-1]]
    for (int i = 0; i < 8; i++) out |= ((~(c >> (7 - i))) & 1u) << i;
    return out;
}

__global__ void k_tx_plan(const c8b_txframe* __restrict__ fr, int n, TxPlan* __restrict__ plan)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const c8b_txframe f = fr[i];
    TxPlan P;
    memset(&P, 0, sizeof P);
    P.format = f.format; P.mcs = f.mcs; P.psdu_len = f.psdu_len; P.psdu_off = f.psdu_off; P.out_off = f.out_off;
    P.cfo_step = (double)f.cfo_hz * 2.0 * 3.14159265358979323846 / 20000000.0;
    int nsym = 0, nslots = 0;
    if (!tx_geometry(f.format, f.mcs, f.psdu_len, &nsym, &nslots)) { P.nslots = 0; plan[i] = P; return; }
    tx_rate(f.format, f.mcs, &P.nbpsc, &P.cr);
    P.nss = tx_nss(f.format, f.mcs);
    P.npre = f.format == C8B_F_L ? 5 : f.format == C8B_F_HT ? 8 + P.nss : 9 + P.nss;
    P.ncbps = (f.format == C8B_F_L ? 48 : 52) * P.nbpsc * P.nss;
    P.ndbps = tx_ndbps(P.ncbps, P.cr);
    P.nsym = nsym; P.nslots = nslots; P.nbits = nsym * P.ndbps;
    P.tail0 = 16 + 8 * f.psdu_len;
    // L-SIG (tools/phy80211.py:236-258): rate, reserved, 12-bit length, even parity, 6 tail
    const uint32_t rateL[8] = { 0xB, 0xF, 0xA, 0xE, 0x9, 0xD, 0x8, 0xC };   // C_LEGACY_RATE_BIT, bit k = element k
    int llen = f.psdu_len;
    if (f.format == C8B_F_HT) llen = ((32 + 4 * P.nss + 4 * nsym - 20) / 4) * 3 - 3;    // txTime = 20 + 8 + 4 + 4 nLTF + 4 nSym
    if (f.format == C8B_F_VHT) llen = ((36 + 4 * P.nss + 4 * nsym - 20) / 4) * 3 - 3;   // + VHT-SIG-B
    uint32_t ls = (f.format == C8B_F_L ? rateL[f.mcs] : rateL[0]) | ((uint32_t)(llen & 0xfff) << 5);
    ls |= (uint32_t)(__popc(ls) & 1) << 17;
    P.lsig = ls;
    if (f.format == C8B_F_HT) {                                             // HT-SIG (tools/phy80211.py:284-330), genFromMpdu: no aggregation
        uint64_t b = (uint64_t)(f.mcs & 0x7f) | ((uint64_t)(f.psdu_len & 0xffff) << 8) | (7ull << 24);
        b |= (uint64_t)crc8_bits(b, 34) << 34;
        P.sig48 = b;
    } else if (f.format == C8B_F_VHT) {                                     // VHT-SIG-A (:355-428): SU, group id 0, partial AID 0, 1 stream
        uint64_t b = (1ull << 2) | ((uint64_t)(P.nss - 1) << 10) | (1ull << 23) | ((uint64_t)(f.mcs & 0xf) << 28) | (1ull << 33);
        b |= (uint64_t)crc8_bits(b, 34) << 34;
        P.sig48 = b;
        // VHT-SIG-B (:520-552): ceil(len / 4) in 17 bits, 3 reserved ones, 6 tail; NDP pattern for an empty A-MPDU
        if (f.psdu_len > 0) {
            const uint32_t sb = (uint32_t)((f.psdu_len + 3) / 4) | (7u << 17);
            P.sigb = sb;
            P.svc_crc = crc8_bits(sb, 20);
        } else {
            P.sigb = 0;                                                     // C_NDP_SIG_B_20, bit k = element k
            const int ndp[20] = { 0, 0, 0, 0, 0, 1, 1, 1, 0, 1, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0 };
            for (int k = 0; k < 20; k++) P.sigb |= (uint32_t)ndp[k] << k;
        }
        // padding (tools/phy80211header.py:478-486)
        const int psdu = (nsym * P.ndbps - 16 - 6) / 8;
        P.npadeof = f.psdu_len > 0 ? (psdu - f.psdu_len) / 4 : 0;
    }
    plan[i] = P;
}

// encoder input bit x of the DATA field (scrambled; tools/phy80211.py:640-708)
__device__ __forceinline__ int data_bit(const TxPlan& P, const uint8_t* __restrict__ psdu, const uint32_t* __restrict__ scr, uint32_t eof, int x)
{
    if (x < 0 || x >= P.nbits) return 0;
    int d = 0;
    if (P.format == C8B_F_VHT) {
        if (x >= P.nbits - 6) return 0;                                     // tail added after scrambling
        if (x >= 8 && x < 16) d = (P.svc_crc >> (x - 8)) & 1;               // service: 7 scrambler init + 1 reserved zeros, SIG-B CRC
        else if (x >= 16) {
            int y = x - 16;
            if (y < 8 * P.psdu_len) d = (psdu[y >> 3] >> (y & 7)) & 1;
            else { y -= 8 * P.psdu_len; if (y < 32 * P.npadeof) d = (eof >> (y & 31)) & 1; }
        }
    } else {
        if (x >= P.tail0 && x < P.tail0 + 6) return 0;                      // tail reset after scrambling
        if (x >= 16 && x < P.tail0) { const int y = x - 16; d = (psdu[y >> 3] >> (y & 7)) & 1; }
    }
    const int s = x % 127;
    return d ^ (int)((scr[s >> 5] >> (s & 31)) & 1u);
}
// output o of the rate-1/2 encoder at step t over a bit source
template <class Src>
__device__ __forceinline__ int bcc_out(Src in, int t, int o)
{
    const int a = in(t) ^ in(t - 2) ^ in(t - 3) ^ in(t - 6);
    return a ^ (o ? in(t - 1) : in(t - 5));                                 // 0155: taps 0,2,3,5,6   0117: taps 0,1,2,3,6
}
// punctured-stream index -> mother-code index (tools/phy80211header.py:773-795)
__device__ __forceinline__ int mother_index(int cr, int u)
{
    if (cr == C8B_CR_12) return u;
    if (cr == C8B_CR_23) return (u / 3) * 4 + u % 3;
    if (cr == C8B_CR_34) { const int r = u & 3; return (u >> 2) * 6 + (r == 3 ? 5 : r); }
    const int q = u / 6, r = u - 6 * q;
    return q * 10 + (r < 3 ? r : r == 3 ? 5 : r == 4 ? 6 : 9);
}
__device__ __forceinline__ float qam_level(int sign, int m)   // Gray-mapped amplitude (C_QAM_MODU_TAB, tools/phy80211header.py:527-575)
{
    return sign ? (float)m : -(float)m;
}

// constellation point of data tone d (0..47 / 0..51) of DATA symbol q in spatial stream iss: the nBPSC coded bits the
// interleaver (and, for two streams of ONE encoder, the stream parser) puts there, Gray-mapped (tools/phy80211.py:712-760)
__device__ __forceinline__ float2 tx_data_tone(const c8b_lut* __restrict__ L, const TxPlan& P, const uint8_t* __restrict__ psdu,
                                               const uint32_t* __restrict__ scr, uint32_t eof, int q, int d, int iss)
{
    const bool leg = P.format == C8B_F_L;
    const int nb = P.nbpsc;
    const int mi = nb == 1 ? 0 : nb == 2 ? 1 : nb == 4 ? 2 : nb == 6 ? 3 : 4;
    int bits = 0;
    auto src = [&](int x) { return data_bit(P, psdu, scr, eof, x); };
    const int sp = nb / 2 > 1 ? nb / 2 : 1;                     // stream parser block (tools/phy80211.py:712-733)
    for (int b = 0; b < nb; b++) {
        const int j = d * nb + b;
        int c = leg ? L->deintL[mi][j] : L->deintNL[iss][mi][j];    // position in this stream's symbol before interleaving
        if (P.nss == 2) c = iss * sp + 2 * sp * (c / sp) + c % sp;  // ... and in the encoder's output
        const int m = mother_index(P.cr, q * P.ncbps + c);
        bits |= bcc_out(src, m >> 1, m & 1) << b;
    }
    const int h = nb >> 1;                                      // bits per axis
    if (nb == 1) return make_float2(bits ? 1.f : -1.f, 0.f);
    const int bi = bits & ((1 << h) - 1), bq = bits >> h;
    auto axis = [&](int a) {
        const int b1 = (a >> 1) & 1, b2 = (a >> 2) & 1, b3 = (a >> 3) & 1;
        int m = 1;
        if (h == 2) m = b1 ? 1 : 3;
        else if (h == 3) m = b1 ? (b2 ? 3 : 1) : (b2 ? 5 : 7);
        else if (h == 4) m = b1 ? (b2 ? (b3 ? 5 : 7) : (b3 ? 3 : 1)) : (b2 ? (b3 ? 11 : 9) : (b3 ? 13 : 15));
        return qam_level(a & 1, m);
    };
    const float nrm = h == 1 ? 0.70710678118654752f : h == 2 ? 0.31622776601683794f : h == 3 ? 0.15430334996209191f : 0.076696498884737041f;
    return make_float2(axis(bi) * nrm, axis(bq) * nrm);
}

// C_STF_L_26 on FFT bin k (+ zeros at +-27, +-28 for HT / VHT)
__device__ __forceinline__ float2 tx_stf(int k)
{
    const int sc = k < 32 ? k : k - 64;
    float r = 0.f;
    if (sc != 0 && (sc & 3) == 0 && sc >= -24 && sc <= 24) {
        const int q = (sc + 24) >> 2;                               // 0..12, 6 = DC
        const int sg[13] = { 1, -1, 1, -1, -1, 1, 0, -1, -1, 1, 1, 1, 1 };
        r = 0.70710678118654752f * (float)sg[q];
    }
    return make_float2(r, r);
}

// the slot's 64 bins (X[g], already synchronised) -> 80 samples: inverse DFT, cyclic prefix / shift, the generator's windowing
// (procConcat2Symbol: boundary samples halved), amplitude and carrier offset
__device__ __forceinline__ void tx_emit(float2 (*X)[64], const float2* tw, int g, int k, int s, int shift, int csd, float gsc, int nslots,
                                        double cfo_step, float2* __restrict__ o)
{
    float ar = 0.f, ai = 0.f;
#pragma unroll 8
    for (int q = 0; q < 64; q++) {
        const float2 x = X[g][q], w = tw[(q * k) & 63];
        ar = fmaf(x.x, w.x, fmaf(-x.y, w.y, ar));
        ai = fmaf(x.x, w.y, fmaf(x.y, w.x, ai));
    }
    __syncthreads();
    X[g][k] = make_float2(ar, ai);
    __syncthreads();
    const bool halfFirst = s == 2 || s >= 4, halfLast = (s == 1 || s >= 3) && s + 1 < nslots;    // procConcat2Symbol
    for (int n = k; n < 80; n += 64) {
        const float2 x = X[g][(n + shift + csd) & 63];
        float a = gsc;
        if ((n == 0 && halfFirst) || (n == 79 && halfLast)) a *= 0.5f;
        float re = x.x * a, im = x.y * a;
        if (cfo_step != 0.0) {                                              // genSignalWithCfo: exp(j i step), i from the first frame sample
            double ph = (double)(s * 80 + n) * cfo_step;
            ph -= 6.283185307179586476925 * floor(ph * 0.15915494309189533577);
            float sn, cs;
            sincosf((float)ph, &sn, &cs);
            const float r2 = re * cs - im * sn;
            im = re * sn + im * cs; re = r2;
        }
        o[n] = make_float2(re, im);
    }
}

__global__ void __launch_bounds__(TXS * 64)
k_tx_slots(const c8b_lut* __restrict__ L, const TxPlan* __restrict__ plan, int nframes, int maxSlots, const uint8_t* __restrict__ psduAll,
           float2* __restrict__ out0, float2* __restrict__ out1, float gain, uint4 scrSeq, uint32_t eof)
{
    const int iss = blockIdx.y;                                             // spatial stream = antenna
    __shared__ float2 X[TXS][64];
    __shared__ float2 tw[64];
    const int g = threadIdx.x >> 6, k = threadIdx.x & 63;
    if (threadIdx.x < 64) tw[k] = make_float2(L->twr[k], -L->twi[k]);      // exp(+2 pi j k / 64)
    const int64_t sid = (int64_t)blockIdx.x * TXS + g;                     // global slot id
    const int f = (int)(sid / maxSlots), s = (int)(sid % maxSlots);
    const bool live = f < nframes && s < plan[f < nframes ? f : 0].nslots && iss < plan[f < nframes ? f : 0].nss;
    float2 v = make_float2(0.f, 0.f);
    float scale = 0.f;
    int shift = 48, csd = 0;
    TxPlan P;
    if (live) {
        P = plan[f];
        const uint32_t scr[4] = { scrSeq.x, scrSeq.y, scrSeq.z, scrSeq.w };
        const uint8_t* __restrict__ psdu = psduAll + P.psdu_off;
        const int nPre = P.npre, nLtf = P.nss;
        const bool pilotBin = (k == 7 || k == 21 || k == 43 || k == 57);
        const float pbase[4] = { 1.f, 1.f, 1.f, -1.f };                     // C_PILOT_L / C_PILOT_HT 1SS / C_PILOT_VHT, subcarriers -21 -7 7 21
        const float pht2[2][4] = { { 1.f, 1.f, -1.f, -1.f }, { 1.f, -1.f, -1.f, 1.f } };   // C_PILOT_HT, two streams
        // cyclic shift of this stream (tools/phy80211header.py:713-725,950-956): 200 ns on the legacy part, 400 ns after it
        if (P.nss == 2 && iss == 1) csd = s < 7 ? 4 : 8;
        const int pslot = k == 43 ? 0 : k == 57 ? 1 : k == 7 ? 2 : 3;
        auto stf = [&]() { return tx_stf(k); };
        if (s < 2) { v = stf(); scale = rsqrtf(12.f); shift = s == 0 ? 32 : 48; }
        else if (s < 4) { v = make_float2(L->ltfL[k], 0.f); scale = rsqrtf(52.f); shift = s == 2 ? 32 : 48; }
        else if (s == 4 || (s < 7 && P.format != C8B_F_L)) {               // L-SIG, HT-SIG 1/2, VHT-SIG-A 1/2: BPSK / QBPSK, 48 tones
            scale = rsqrtf(52.f);
            const int c = L->sigDemap[k];
            if (pilotBin) v = make_float2(pbase[pslot], 0.f);
            else if (c >= 0) {
                const uint64_t bits = s == 4 ? (uint64_t)P.lsig : P.sig48;
                const int cc = c + (s == 6 ? 48 : 0);
                auto src = [&](int x) { return x < 0 ? 0 : (int)((bits >> x) & 1ull); };
                const float a = bcc_out(src, cc >> 1, cc & 1) ? 1.f : -1.f;
                const bool q = (P.format == C8B_F_HT && s >= 5) || (P.format == C8B_F_VHT && s == 6);   // QBPSK
                v = q ? make_float2(0.f, a) : make_float2(a, 0.f);
            }
        } else if (s == 7 && P.format != C8B_F_L) { v = stf(); scale = rsqrtf(12.f); }   // HT-STF / VHT-STF
        else if (s >= 8 && s < 8 + nLtf && P.format != C8B_F_L) {            // HT-LTF / VHT-LTF l: stream sign P[iss][l], VHT pilot tones R[l]
            const int l = s - 8;
            float sg = (l == 1 && iss == 0) ? -1.f : 1.f;                   // C_P_LTF_VHT_4 rows 0 / 1, columns 0 / 1
            if (P.format == C8B_F_VHT && pilotBin) sg = l == 1 ? -1.f : 1.f; // C_R_LTF_VHT_4
            v = make_float2(L->ltfNL[k] * sg, 0.f);
            scale = rsqrtf(56.f);
        } else if (s == 8 + nLtf && P.format == C8B_F_VHT) {               // VHT-SIG-B: 26 bits, BPSK on 52 tones, pilots without polarity
            scale = rsqrtf(56.f);
            const int d = L->binToDataNL[k];
            if (pilotBin) v = make_float2(pbase[pslot], 0.f);
            else if (d != 255) {
                const int c = L->deintNL[0][0][d];
                const uint32_t bits = P.sigb;
                auto src = [&](int x) { return x < 0 ? 0 : (int)((bits >> x) & 1u); };
                v = make_float2(bcc_out(src, c >> 1, c & 1) ? 1.f : -1.f, 0.f);
            }
        } else {                                                            // DATA symbol q
            const int q = s - nPre;
            const bool leg = P.format == C8B_F_L;
            scale = rsqrtf(leg ? 52.f : 56.f);
            const int d = leg ? L->binToDataL[k] : L->binToDataNL[k];
            if (pilotBin) {
                const int idx0 = leg ? 1 : P.format == C8B_F_HT ? 3 : 4;    // tools/phy80211.py:786-801
                const float pol = L->pilotP[(idx0 + q) % 127];
                const float* pb4 = (P.format == C8B_F_HT && P.nss == 2) ? pht2[iss] : pbase;
                v = make_float2(pol * (leg ? pbase[pslot] : pb4[(pslot + q) & 3]), 0.f);
            } else if (d != 255) {
                v = tx_data_tone(L, P, psdu, scr, eof, q, d, iss);
            }
        }
    }
    X[g][k] = v;
    __syncthreads();
    if (!live) return;                                                      // (threads that have exited do not count at the barriers below)
    const float gsc = gain * scale * (1.0f / 64.0f) * (P.nss == 2 ? 0.70710678118654752f : 1.0f);       // procToneScaling: / sqrt(N_tone nSS)
    tx_emit(X, tw, g, k, s, shift, csd, gsc, P.nslots, P.cfo_step, (iss ? out1 : out0) + P.out_off + (int64_t)s * 80);
}

// ---------------------------------------------------------------------------------------------------------------------
// Two-user VHT MU-MIMO (genAmpduMu, tools/phy80211.py:180-221): one space-time stream per user, each with its own A-MPDU,
// MCS, VHT-SIG-B and padding to the common symbol count; from the VHT-STF on, the two streams (stream 1 cyclically shifted by
// 400 ns) go through the per-subcarrier spatial mapping matrix Q (procSpatialMapping: antenna t = sum_u Q[k][t][u] X_u[k]).
// The legacy part is the two-antenna one (200 ns shift on antenna 1).  plan[2 f + u] = user u of frame f.
// ---------------------------------------------------------------------------------------------------------------------
__host__ __device__ inline int tx_mu_geometry(int mcs0, int len0, int mcs1, int len1, int* nsym, int* nslots)
{
    int ns = 0;
    const int mcs[2] = { mcs0, mcs1 }, len[2] = { len0, len1 };
    for (int u = 0; u < 2; u++) {
        int nbpsc = 0, cr = 0;
        if (!tx_rate(C8B_F_VHT, mcs[u], &nbpsc, &cr) || len[u] < 4 || len[u] > 4095 || (len[u] & 3)) return 0;
        const int ndbps = tx_ndbps(52 * nbpsc, cr);
        const int n = (8 * len[u] + 16 + 6 + ndbps - 1) / ndbps;
        if (n > ns) ns = n;
    }
    *nsym = ns;
    *nslots = 11 + ns;                                                      // L-STF 2, L-LTF 2, L-SIG, SIG-A 2, VHT-STF, 2 VHT-LTF, SIG-B
    return 1;
}

__global__ void k_tx_plan_mu(const c8b_txmu* __restrict__ fr, int n, TxPlan* __restrict__ plan)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const c8b_txmu f = fr[i];
    int nsym = 0, nslots = 0;
    const bool ok = tx_mu_geometry(f.mcs[0], f.psdu_len[0], f.mcs[1], f.psdu_len[1], &nsym, &nslots) != 0;
    // L-SIG: txTime = 20 + 8 + 4 + 4 nLTF + 4 + 4 nSym with two VHT-LTFs (procPktLenAggreMu, tools/phy80211header.py:509-525)
    const int llen = ((36 + 4 * 2 + 4 * nsym - 20) / 4) * 3 - 3;
    uint32_t ls = 0xBu | ((uint32_t)(llen & 0xfff) << 5);
    ls |= (uint32_t)(__popc(ls) & 1) << 17;
    // VHT-SIG-A, MU form (tools/phy80211.py:355-428): group id, one space-time stream for users 0 and 1, BCC everywhere
    uint64_t a = (1ull << 2) | ((uint64_t)(f.group_id & 63) << 4) | (1ull << 10) | (1ull << 13) | (1ull << 23) | (0x1full << 29);
    a |= (uint64_t)crc8_bits(a, 34) << 34;
    for (int u = 0; u < 2; u++) {
        TxPlan P;
        memset(&P, 0, sizeof P);
        P.format = C8B_F_VHT; P.mcs = f.mcs[u]; P.psdu_len = f.psdu_len[u]; P.psdu_off = f.psdu_off[u]; P.out_off = f.out_off;
        P.cfo_step = (double)f.cfo_hz * 2.0 * 3.14159265358979323846 / 20000000.0;
        P.q_off = f.q_index * 256;
        if (ok) {
            tx_rate(C8B_F_VHT, f.mcs[u], &P.nbpsc, &P.cr);
            P.nss = 1; P.npre = 11;
            P.ncbps = 52 * P.nbpsc;
            P.ndbps = tx_ndbps(P.ncbps, P.cr);
            P.nsym = nsym; P.nslots = nslots; P.nbits = nsym * P.ndbps;
            P.lsig = ls; P.sig48 = a;
            // VHT-SIG-B, MU form (:571-600): floor(len / 4) in 16 bits, MCS in 4, 6 tail; its CRC rides in the service field
            const uint32_t sb = (uint32_t)(f.psdu_len[u] / 4) | ((uint32_t)(f.mcs[u] & 0xf) << 16);
            P.sigb = sb;
            P.svc_crc = crc8_bits(sb, 20);
            const int psdu = (nsym * P.ndbps - 16 - 6) / 8;
            P.npadeof = (psdu - f.psdu_len[u]) / 4;
        }
        plan[2 * i + u] = P;
    }
}

__global__ void __launch_bounds__(TXS * 64)
k_tx_slots_mu(const c8b_lut* __restrict__ L, const TxPlan* __restrict__ plan, int nframes, int maxSlots, const uint8_t* __restrict__ psduAll,
              const float2* __restrict__ Q, float2* __restrict__ out0, float2* __restrict__ out1, float gain, uint4 scrSeq, uint32_t eof)
{
    const int t = blockIdx.y;                                               // transmit antenna
    __shared__ float2 X[TXS][64];
    __shared__ float2 tw[64];
    const int g = threadIdx.x >> 6, k = threadIdx.x & 63;
    if (threadIdx.x < 64) tw[k] = make_float2(L->twr[k], -L->twi[k]);      // exp(+2 pi j k / 64)
    __syncthreads();
    const int64_t sid = (int64_t)blockIdx.x * TXS + g;
    const int f = (int)(sid / maxSlots), s = (int)(sid % maxSlots);
    const bool live = f < nframes && s < plan[2 * (f < nframes ? f : 0)].nslots;
    float2 v = make_float2(0.f, 0.f);
    float scale = 0.f;
    int shift = 48, csd = 0, nslots = 0;
    double cfo = 0.0;
    int64_t outOff = 0;
    if (live) {
        const TxPlan P0 = plan[2 * f];
        nslots = P0.nslots; cfo = P0.cfo_step; outOff = P0.out_off;
        const uint32_t scr[4] = { scrSeq.x, scrSeq.y, scrSeq.z, scrSeq.w };
        const bool pilotBin = (k == 7 || k == 21 || k == 43 || k == 57);
        const float pbase[4] = { 1.f, 1.f, 1.f, -1.f };
        const int pslot = k == 43 ? 0 : k == 57 ? 1 : k == 7 ? 2 : 3;
        if (s < 7) {                                                        // legacy part: as any two-antenna frame
            if (t == 1) csd = 4;
            if (s < 2) { v = tx_stf(k); scale = rsqrtf(12.f); shift = s == 0 ? 32 : 48; }
            else if (s < 4) { v = make_float2(L->ltfL[k], 0.f); scale = rsqrtf(52.f); shift = s == 2 ? 32 : 48; }
            else {
                scale = rsqrtf(52.f);
                const int c = L->sigDemap[k];
                if (pilotBin) v = make_float2(pbase[pslot], 0.f);
                else if (c >= 0) {
                    const uint64_t bits = s == 4 ? (uint64_t)P0.lsig : P0.sig48;
                    const int cc = c + (s == 6 ? 48 : 0);
                    auto src = [&](int x) { return x < 0 ? 0 : (int)((bits >> x) & 1ull); };
                    const float a = bcc_out(src, cc >> 1, cc & 1) ? 1.f : -1.f;
                    v = s == 6 ? make_float2(0.f, a) : make_float2(a, 0.f);
                }
            }
        } else {                                                            // VHT part: X_0, X_1 of this bin, then row t of Q
            float2 xu[2];
            scale = s == 7 ? rsqrtf(12.f) : rsqrtf(56.f);
#pragma unroll
            for (int u = 0; u < 2; u++) {
                const TxPlan& P = u ? plan[2 * f + 1] : P0;
                float2 x = make_float2(0.f, 0.f);
                if (s == 7) x = tx_stf(k);
                else if (s < 10) {                                          // VHT-LTF l: P[u][l] on data tones, R[l] on pilot tones
                    const int l = s - 8;
                    const float sg = pilotBin ? (l == 1 ? -1.f : 1.f) : ((l == 1 && u == 0) ? -1.f : 1.f);
                    x = make_float2(L->ltfNL[k] * sg, 0.f);
                } else if (s == 10) {                                       // the user's own VHT-SIG-B
                    const int d = L->binToDataNL[k];
                    if (pilotBin) x = make_float2(pbase[pslot], 0.f);
                    else if (d != 255) {
                        const int c = L->deintNL[0][0][d];
                        const uint32_t bits = P.sigb;
                        auto src = [&](int y) { return y < 0 ? 0 : (int)((bits >> y) & 1u); };
                        x = make_float2(bcc_out(src, c >> 1, c & 1) ? 1.f : -1.f, 0.f);
                    }
                } else {
                    const int q = s - 11;
                    const int d = L->binToDataNL[k];
                    if (pilotBin) x = make_float2(L->pilotP[(4 + q) % 127] * pbase[(pslot + q) & 3], 0.f);
                    else if (d != 255) x = tx_data_tone(L, P, psduAll + P.psdu_off, scr, eof, q, d, 0);     // each user is a one-stream frame
                }
                xu[u] = x;
            }
            const float2 w = tw[(8 * k) & 63];                              // stream 1: 400 ns = 8 samples (C_CYCLIC_SHIFT_NL)
            xu[1] = make_float2(xu[1].x * w.x - xu[1].y * w.y, xu[1].x * w.y + xu[1].y * w.x);
            const float2* __restrict__ q = Q + P0.q_off + 4 * ((k + 32) & 63) + 2 * t;     // Q[natural index][t][u]
            const float2 q0 = q[0], q1 = q[1];
            v = make_float2(q0.x * xu[0].x - q0.y * xu[0].y + q1.x * xu[1].x - q1.y * xu[1].y,
                            q0.x * xu[0].y + q0.y * xu[0].x + q1.x * xu[1].y + q1.y * xu[1].x);
        }
    }
    X[g][k] = v;
    __syncthreads();
    if (!live) return;
    const float gsc = gain * scale * (1.0f / 64.0f) * 0.70710678118654752f;
    tx_emit(X, tw, g, k, s, shift, csd, gsc, nslots, cfo, (t ? out1 : out0) + outOff + (int64_t)s * 80);
}

// Synthetic traffic: every frame gets its own random MPDU with a valid FCS (CRC-32 as tools/mac80211.py:36-47 appends it),
// wrapped for VHT in a one-MPDU A-MPDU (delimiter of tools/mac80211.py:333-360: EOF, reserved, len[12:14], len[0:12], CRC-8,
// 0x4E) -- so a decoded PDU can be compared byte for byte with what was sent.  One thread per frame.
__device__ __forceinline__ uint64_t mix64(uint64_t z)
{
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
__global__ void k_tx_fill(const c8b_lut* __restrict__ L, const c8b_txframe* __restrict__ fr, int n, uint8_t* __restrict__ psduAll, uint64_t seed)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const c8b_txframe f = fr[i];
    uint8_t* __restrict__ p = psduAll + f.psdu_off;
    int mlen = f.psdu_len;
    if (f.format == C8B_F_VHT) {
        mlen = f.psdu_len - 4;
        const uint32_t d = 1u | (((uint32_t)mlen >> 12) & 3u) << 2 | ((uint32_t)mlen & 0xfffu) << 4;
        const uint32_t c = crc8_bits(d, 16);
        p[0] = (uint8_t)d; p[1] = (uint8_t)(d >> 8); p[2] = (uint8_t)c; p[3] = 0x4E;
        p += 4;
    }
    if (mlen < 4) return;
    uint32_t crc = 0xffffffffu;
    for (int k = 0; k < mlen - 4; k++) {
        const uint64_t r = mix64(seed ^ ((uint64_t)i << 24) ^ (uint64_t)(k >> 3));
        const uint8_t b = (uint8_t)(r >> (8 * (k & 7)));
        p[k] = b;
        crc = L->crc32tab[(crc ^ b) & 0xff] ^ (crc >> 8);
    }
    crc = ~crc;
    for (int k = 0; k < 4; k++) p[mlen - 4 + k] = (uint8_t)(crc >> (8 * k));
}

}  // namespace

void c8b_launch_tx_fill(const c8b_lut* lut, const c8b_txframe* d_frames, int nframes, uint8_t* d_psdu, uint64_t seed, cudaStream_t st)
{
    if (nframes <= 0) return;
    k_tx_fill<<<(nframes + 127) / 128, 128, 0, st>>>(lut, d_frames, nframes, d_psdu, seed);
}

// C_VHT_EOF (tools/phy80211header.py:737): EOF delimiter of length 0, bit k = element k
uint32_t c8b_tx_eof_word(void) { return 1u | (crc8_bits(1u, 16) << 16) | (0x4Eu << 24); }

// scrambler sequence of procScramble (tools/phy80211header.py:800-806) for a 7-bit seed: 127 bits, bit i = feedback of step i
void c8b_tx_scrambler(int seed, uint32_t out[4])
{
    out[0] = out[1] = out[2] = out[3] = 0;
    int st = seed & 0x7f;
    for (int i = 0; i < 127; i++) {
        const int fb = ((st >> 6) ^ (st >> 3)) & 1;
        out[i >> 5] |= (uint32_t)fb << (i & 31);
        st = ((st << 1) & 0x7e) | fb;
    }
}

int c8b_tx_geometry_host(int format, int mcs, int len, int* nsym, int* nslots) { return tx_geometry(format, mcs, len, nsym, nslots); }

size_t c8b_tx_plan_bytes(int nframes) { return (size_t)nframes * sizeof(TxPlan); }

int c8b_tx_nss_host(int format, int mcs) { return tx_nss(format, mcs); }

int c8b_tx_mu_geometry_host(int mcs0, int len0, int mcs1, int len1, int* nsym, int* nslots) { return tx_mu_geometry(mcs0, len0, mcs1, len1, nsym, nslots); }

void c8b_launch_tx_mu(const c8b_lut* lut, const c8b_txmu* d_frames, int nframes, int maxSlots, void* d_plan, const uint8_t* d_psdu, const float2* d_q,
                      float2* d_out0, float2* d_out1, float gain, const uint32_t scr[4], uint32_t eof, cudaStream_t st)
{
    if (nframes <= 0 || maxSlots <= 0) return;
    TxPlan* plan = reinterpret_cast<TxPlan*>(d_plan);                       // two entries per frame
    k_tx_plan_mu<<<(nframes + 127) / 128, 128, 0, st>>>(d_frames, nframes, plan);
    const int64_t slots = (int64_t)nframes * maxSlots;
    const dim3 grid((unsigned)((slots + TXS - 1) / TXS), 2);                // y = transmit antenna
    k_tx_slots_mu<<<grid, TXS * 64, 0, st>>>(lut, plan, nframes, maxSlots, d_psdu, d_q, d_out0, d_out1, gain, make_uint4(scr[0], scr[1], scr[2], scr[3]), eof);
}

void c8b_launch_tx(const c8b_lut* lut, const c8b_txframe* d_frames, int nframes, int maxSlots, void* d_plan, const uint8_t* d_psdu,
                   float2* d_out, float2* d_out1, float gain, const uint32_t scr[4], uint32_t eof, cudaStream_t st)
{
    if (nframes <= 0 || maxSlots <= 0) return;
    TxPlan* plan = reinterpret_cast<TxPlan*>(d_plan);
    k_tx_plan<<<(nframes + 127) / 128, 128, 0, st>>>(d_frames, nframes, plan);
    const int64_t slots = (int64_t)nframes * maxSlots;
    const dim3 grid((unsigned)((slots + TXS - 1) / TXS), d_out1 ? 2 : 1);                   // y = spatial stream
    k_tx_slots<<<grid, TXS * 64, 0, st>>>(lut, plan, nframes, maxSlots, d_psdu, d_out, d_out1, gain, make_uint4(scr[0], scr[1], scr[2], scr[3]), eof);
}
