// TEST INFRASTRUCTURE (oracle/_ref): thin extern "C" doors onto the UNMODIFIED reference
// translation unit /root/reference/lib/cloud80211phy.cc (compiled where it lies, see Makefile).
// Nothing here restates an algorithm: every function forwards to the reference symbol named
// in its comment.  Used (a) to pin oracle/oracle_rx.cc, (b) to generate tests/golden/*.npz
// (tests/golden/make_golden.py), (c) as the per-function CPU baseline ("kind":"reference").
// Never linked into the product library.
#include "cloud80211phy.h"
#include <cstring>

extern "C" {

// cloud80211phy.h:58-98 -- c8p_mod is 65 plain ints; exported as a flat int[65].
int ref_mod_nints(void) { return (int)(sizeof(c8p_mod) / sizeof(int)); }

// cloud80211phy.cc:609-627
void ref_lsig_demod(const float* sym1, const float* sym2, const float* sig, float* h_out, float* llr48)
{
    std::vector<gr_complex> h(64, gr_complex(0.0f, 0.0f));
    gr_complex a[64], b[64], c[64];
    memcpy(a, sym1, sizeof(a)); memcpy(b, sym2, sizeof(b)); memcpy(c, sig, sizeof(c));
    procLHSigDemodDeint(a, b, c, h, llr48);
    memcpy(h_out, h.data(), sizeof(gr_complex) * 64);
}

// cloud80211phy.cc:629-648
void ref_nlsig_demod(const float* sym1, const float* sym2, const float* h_in, float* llrht96, float* llrvht96)
{
    std::vector<gr_complex> h(64);
    memcpy(h.data(), h_in, sizeof(gr_complex) * 64);
    gr_complex a[64], b[64];
    memcpy(a, sym1, sizeof(a)); memcpy(b, sym2, sizeof(b));
    procNLSigDemodDeint(a, b, h, llrht96, llrvht96);
}

// cloud80211phy.cc:2001-2088 (svSigDecoder::decode, trellisLen <= 48)
void ref_sig_viterbi(const float* llr, uint8_t* bits, int trellisLen)
{
    static thread_local svSigDecoder dec;
    std::vector<float> tmp(llr, llr + 2 * trellisLen);
    dec.decode(tmp.data(), bits, trellisLen);
}

// cloud80211phy.cc:1890-1999 (SV_Decode_Sig, any length, rate 1/2 input = 2 LLR per step)
void ref_sv_decode(const float* llr, uint8_t* bits, int trellisLen)
{
    std::vector<float> tmp(llr, llr + 2 * (size_t)trellisLen);
    SV_Decode_Sig(tmp.data(), bits, trellisLen);
}

// cloud80211phy.cc:650-728
int ref_check_legacy(const uint8_t* bits, int* mcs, int* len, int* ndbps)
{
    uint8_t b[24]; memcpy(b, bits, 24);
    return signalCheckLegacy(b, mcs, len, ndbps) ? 1 : 0;
}
// cloud80211phy.cc:730-751
int ref_check_ht(const uint8_t* bits) { uint8_t b[48]; memcpy(b, bits, 48); return signalCheckHt(b) ? 1 : 0; }
// cloud80211phy.cc:753-771
int ref_check_vhta(const uint8_t* bits) { uint8_t b[48]; memcpy(b, bits, 48); return signalCheckVhtA(b) ? 1 : 0; }
// cloud80211phy.cc:1367-1403
int ref_crc8_check(const uint8_t* bits, int len, const uint8_t* crc)
{
    std::vector<uint8_t> b(bits, bits + len); uint8_t c[8]; memcpy(c, crc, 8);
    return checkBitCrc8(b.data(), len, c) ? 1 : 0;
}
// cloud80211phy.cc:1325-1365
void ref_crc8_gen(const uint8_t* bits, int len, uint8_t* crc)
{
    std::vector<uint8_t> b(bits, bits + len);
    genCrc8Bits(b.data(), crc, len);
}

// cloud80211phy.cc:773-850
void ref_parse_l(int mcs, int len, int* mod65)
{
    c8p_mod m; memset(&m, 0, sizeof(m));
    signalParserL(mcs, len, &m);
    memcpy(mod65, &m, sizeof(m));
}
// cloud80211phy.cc:852-998
void ref_parse_ht(const uint8_t* bits, int* mod65)
{
    c8p_mod m; c8p_sigHt s; memset(&m, 0, sizeof(m));
    uint8_t b[48]; memcpy(b, bits, 48);
    signalParserHt(b, &m, &s);
    memcpy(mod65, &m, sizeof(m));
}
// cloud80211phy.cc:1090-1178
void ref_parse_vhta(const uint8_t* bits, int* mod65)
{
    c8p_mod m; c8p_sigVhtA s; memset(&m, 0, sizeof(m));
    uint8_t b[48]; memcpy(b, bits, 48);
    signalParserVhtA(b, &m, &s);
    memcpy(mod65, &m, sizeof(m));
}
// cloud80211phy.cc:1180-1223 (in/out: mod65 carries the state left by ref_parse_vhta)
void ref_parse_vhtb(const uint8_t* bits, int* mod65)
{
    c8p_mod m; memcpy(&m, mod65, sizeof(m));
    uint8_t b[26]; memcpy(b, bits, 26);
    signalParserVhtB(b, &m);
    memcpy(mod65, &m, sizeof(m));
}
// cloud80211phy.cc:1225-1323 (mod65[nSS] must be set by the caller, as demod does via VHT-SIG-A)
void ref_mod_vht(int mcs, int nss, int* mod65)
{
    c8p_mod m; memset(&m, 0, sizeof(m)); m.nSS = nss;
    modParserVht(mcs, &m);
    memcpy(mod65, &m, sizeof(m));
}
// cloud80211phy.cc:1000-1088
void ref_mod_ht(int mcs, int nss, int* mod65)
{
    c8p_mod m; memset(&m, 0, sizeof(m)); m.nSS = nss;
    modParserHt(mcs, &m);
    memcpy(mod65, &m, sizeof(m));
}

// cloud80211phy.cc:2090-2148 ; qam is nSD complex (modified in place by the reference), llr is nSD*nBPSCS
void ref_qam_to_llr(float* qam, float* llr, int mod, int nSD)
{
    c8p_mod m; memset(&m, 0, sizeof(m)); m.mod = mod; m.nSD = nSD;
    procSymQamToLlr(reinterpret_cast<gr_complex*>(qam), llr, &m);
}
// cloud80211phy.cc:2150-2192
void ref_deint_l(const float* in, float* out, int nCBPS)
{
    c8p_mod m; memset(&m, 0, sizeof(m)); m.nCBPS = nCBPS;
    std::vector<float> t(in, in + nCBPS);
    procSymDeintL2(t.data(), out, &m);
}
// cloud80211phy.cc:2238-2287 / 2289-2338 ; ss = 1 or 2
void ref_deint_nl(const float* in, float* out, int nCBPSS, int ss)
{
    c8p_mod m; memset(&m, 0, sizeof(m)); m.nCBPSS = nCBPSS;
    std::vector<float> t(in, in + nCBPSS);
    if (ss == 1) procSymDeintNL2SS1(t.data(), out, &m);
    else procSymDeintNL2SS2(t.data(), out, &m);
}
// cloud80211phy.cc:2442-2451
void ref_depas(const float* in0, const float* in1, float* out, int nBPSCS, int nCBPSS)
{
    static thread_local float in[C8P_MAX_N_SS][C8P_MAX_N_CBPSS];
    memcpy(in[0], in0, sizeof(float) * nCBPSS); memcpy(in[1], in1, sizeof(float) * nCBPSS);
    c8p_mod m; memset(&m, 0, sizeof(m)); m.nBPSCS = nBPSCS; m.nCBPSS = nCBPSS;
    procSymDepasNL(in, out, &m);
}
// cloud80211phy.cc:2622-2644
void ref_bcc(const uint8_t* in, uint8_t* out, int len)
{
    std::vector<uint8_t> t(in, in + len);
    bccEncoder(t.data(), out, len);
}
// cloud80211phy.cc:1849-1855
void ref_intl_vhtb20(const uint8_t* in, uint8_t* out) { uint8_t t[52]; memcpy(t, in, 52); procIntelVhtB20(t, out); }
// cloud80211phy.cc:2594-2606 (TX scrambler, used by tests to cross-check the RX descrambler)
void ref_scramble(const uint8_t* in, uint8_t* out, int len, int init)
{
    std::vector<uint8_t> t(in, in + len);
    scramEncoder(t.data(), out, len, init);
}

// extern const tables of cloud80211phy.h:151-195; returns element count, copies as int32/float32
int ref_table_i(const char* name, int* out)
{
#define TI(N, L) if (!strcmp(name, #N)) { for (int i = 0; i < (L); i++) out[i] = (int)(N)[i]; return (L); }
    TI(mapDeintLegacyBpsk, 48) TI(mapDeintLegacyQpsk, 96) TI(mapDeintLegacy16Qam, 192) TI(mapDeintLegacy64Qam, 288)
    TI(mapDeintNonlegacyBpsk, 52) TI(mapDeintNonlegacyQpsk, 104) TI(mapDeintNonlegacy16Qam, 208)
    TI(mapDeintNonlegacy64Qam, 312) TI(mapDeintNonlegacy256Qam, 416) TI(mapDeintVhtSigB20, 52)
    TI(SV_PUNC_12, 2) TI(SV_PUNC_23, 4) TI(SV_PUNC_34, 6) TI(SV_PUNC_56, 10)
    TI(FFT_26_DEMAP, 64) TI(C8P_LEGACY_D_SC, 64) TI(EOF_PAD_SUBFRAME, 32)
    if (!strcmp(name, "SV_STATE_NEXT")) { for (int i = 0; i < 128; i++) out[i] = SV_STATE_NEXT[i / 2][i % 2]; return 128; }
    if (!strcmp(name, "SV_STATE_OUTPUT")) { for (int i = 0; i < 128; i++) out[i] = SV_STATE_OUTPUT[i / 2][i % 2]; return 128; }
#undef TI
    return -1;
}
int ref_table_f(const char* name, float* out)
{
#define TF(N, L) if (!strcmp(name, #N)) { for (int i = 0; i < (L); i++) out[i] = (N)[i]; return (L); }
    TF(LTF_L_26_F_FLOAT, 64) TF(LTF_NL_28_F_FLOAT, 64) TF(LTF_NL_28_F_FLOAT_VHT22, 64) TF(LTF_NL_28_F_FLOAT2, 64)
    TF(PILOT_P, 127) TF(PILOT_L, 4) TF(PILOT_HT_1, 4) TF(PILOT_HT_2_1, 4) TF(PILOT_HT_2_2, 4) TF(PILOT_VHT, 4)
#undef TF
    return -1;
}

} // extern "C"
