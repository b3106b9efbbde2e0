/* rx_impl.cc -- gr::ieee80211::rx (include/gnuradio/ieee80211/rx.h): the receive chain as one sink block over the
 * live-stream session of libc80211b200.so.  Frames, tags-as-fields and PDUs are the ones the seven-block chain publishes
 * (tests/test_gpu_stream.py pins the session to one whole-capture pass and to the oracle for any call sizes;
 * tests/test_gr_shells.py runs this block under the mock runtime).
 */
#include <gnuradio/ieee80211/rx.h>
#include <gnuradio/io_signature.h>

#include <c80211b200.h>

#include <algorithm>
#include <iostream>
#include <stdexcept>
#include <string>
#include <vector>

namespace gr {
namespace ieee80211 {

namespace {

constexpr int kPduStride = 4400;     // one frame's PDU records: <= 4095 + 4 bytes for one MPDU, A-MPDU subframes share the PSDU

class rx_impl : public rx
{
    c8b_ctx* d_ctx = nullptr;
    const int d_nant;
    const bool d_debug;
    std::vector<c8b_frame> d_frames;
    std::vector<int64_t> d_base;
    std::vector<uint8_t> d_pdu;
    long d_nPktCorrect = 0, d_legacyMcsCount[8] = { 0 }, d_htMcsCount[8] = { 0 }, d_vhtMcsCount[10] = { 0 };

public:
    rx_impl(int nant, int mupos, int mugid, bool debug)
        : gr::block("rx", gr::io_signature::make(nant, nant, sizeof(gr_complex)), gr::io_signature::make(0, 0, 0)), d_nant(nant), d_debug(debug)
    {
        if (nant != 1 && nant != 2) throw std::invalid_argument("ieee80211 rx: nant must be 1 or 2");
        c8b_cfg cfg = c8b_cfg();
        cfg.max_frames = 64;                                       // frame records per window pass
        cfg.chunk_items = 1;
        cfg.mupos = mupos;
        cfg.mugid = mugid;
        if (c8b_create(&cfg, &d_ctx) != C8B_OK) throw std::runtime_error(std::string("ieee80211 rx: ") + c8b_last_error(nullptr));
        std::vector<uint8_t> lut(c8b_lut_size());
        if (c8b_lut_blob(lut.data(), lut.size()) || c8b_lut_load(d_ctx, lut.data(), lut.size()) || c8b_stream_begin(d_ctx, nant, 0)) {
            const std::string why = c8b_last_error(d_ctx);
            c8b_destroy(d_ctx);
            throw std::runtime_error("ieee80211 rx: " + why);
        }
        message_port_register_out(pmt::mp("out"));
        set_output_multiple(4096);                                 // fewer, larger calls
    }
    ~rx_impl() override { c8b_destroy(d_ctx); }

    void forecast(int noutput_items, gr_vector_int& ninput_items_required) override
    {
        for (auto& r : ninput_items_required) r = noutput_items;
    }

    int general_work(int, gr_vector_int& ninput_items, gr_vector_const_void_star& input_items, gr_vector_void_star&) override
    {
        const int n = *std::min_element(ninput_items.begin(), ninput_items.end());
        push(static_cast<const float*>(input_items[0]), d_nant == 2 ? static_cast<const float*>(input_items[1]) : nullptr, n, 0);
        consume_each(n);
        return 0;
    }

    bool stop() override                                           // end of the flowgraph: what is left is decided now
    {
        push(nullptr, nullptr, 0, 1);
        return true;
    }

private:
    void push(const float* a, const float* b, int n, int flush)
    {
        const int cap = n / 400 + 4 * 64 + 64;                     // more than one push can release (c80211b200.h)
        if ((int)d_frames.size() < cap) { d_frames.resize(cap); d_base.resize(cap); d_pdu.resize((size_t)cap * kPduStride); }
        int nf = 0;
        const int rc = c8b_stream_push(d_ctx, a, b, n, flush, d_frames.data(), cap, &nf, d_base.data(), d_pdu.data(), kPduStride);
        if (rc != C8B_OK) {                                        // never throw from work(): report and resynchronise
            std::cout << "ieee80211 rx, error: " << c8b_last_error(d_ctx) << std::endl;
            if (rc == C8B_ERR_FULL) c8b_stream_begin(d_ctx, d_nant, 0);
            return;
        }
        for (int k = 0; k < nf; k++) {
            const c8b_frame& f = d_frames[k];
            const uint8_t* p = d_pdu.data() + (size_t)k * kPduStride;
            for (int q = 0, o = 0; q < f.npdu && o + 3 <= f.pdu_bytes; q++) {
                const int len = p[o + 1] | (p[o + 2] << 8);
                const int m = p[o] == 20 ? len + 3 : len + 4;      // NDP channel report : [fmt][len][MPDU][mcs]
                pmt::pmt_t meta = pmt::dict_add(pmt::make_dict(), pmt::mp("len"), pmt::from_long(m));
                message_port_pub(pmt::mp("out"), pmt::cons(meta, pmt::make_blob(p + o, m)));
                o += m;
            }
            if (d_debug && f.status == C8B_ST_OK) debug_line(f);
        }
    }

    // decode(ifdebug = true)'s lines (lib/decode_impl.cc:377-411,456-509), which tools/performance/perf_siso.py scrapes
    void debug_line(const c8b_frame& f)
    {
        const bool ok = f.npdu > 0;
        if (!ok && f.format == C8B_F_HT && f.ampdu) return;
        long* cnt = f.format == C8B_F_VHT ? d_vhtMcsCount : f.format == C8B_F_HT ? d_htMcsCount : d_legacyMcsCount;
        const int ncnt = f.format == C8B_F_VHT ? 10 : 8;
        for (int k = 0; k < std::max(f.npdu, 1); k++) {
            if (ok) {
                d_nPktCorrect++;
                if (f.format == C8B_F_VHT) { if (f.mcs >= 0 && f.mcs < 10) cnt[f.mcs]++; }
                else cnt[((f.mcs % 8) + 8) % 8]++;
            }
            std::string s = std::string("ieee80211 decode, ") + (f.format == C8B_F_VHT ? "vht" : f.format == C8B_F_HT ? "ht" : "legacy") +
                            " crc32 " + (ok ? "correct" : "wrong") + ", total:" + std::to_string(d_nPktCorrect);
            for (int i = 0; i < ncnt; i++) s += "," + std::to_string(i) + ":" + std::to_string(cnt[i]);
            s += ",cfo:" + std::to_string(f.cfo_hz) + ",snr:" + std::to_string(f.snr) + ",rssi:" + std::to_string(f.rssi);
            if (f.format == C8B_F_VHT) s += ",sssnr0:" + std::to_string(f.sssnr0) + ",sssnr1:" + std::to_string(f.sssnr1);
            std::cout << s << std::endl;
        }
    }
};

}  // namespace

rx::sptr rx::make(int nant, int mupos, int mugid, bool ifdebug) { return gnuradio::make_block_sptr<rx_impl>(nant, mupos, mugid, ifdebug); }

}  // namespace ieee80211
}  // namespace gr
