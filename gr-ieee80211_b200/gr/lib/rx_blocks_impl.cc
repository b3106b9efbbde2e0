/* rx_blocks_impl.cc -- the seven receive blocks of gr-ieee80211 as shells around libc80211b200.so.
 *
 * Drop-in for lib/{trigger,sync,signal,signal2,demod,demod2,decode}_impl.{cc,h} of the reference: it defines the same
 * gr::ieee80211::X::make() the public headers include/gnuradio/ieee80211/X.h declare, so the pybind bindings
 * (python/ieee80211/bindings/X_python.cc), the GRC YAML (grc/ieee80211_X.block.yml) and every .grc flowgraph stay as they
 * are.  In lib/CMakeLists.txt replace the seven *_impl.cc by this file and add `c80211b200` to target_link_libraries
 * (gr/README.md).
 *
 * A shell owns one c8b_blk (its own CUDA context and stream: GNU Radio runs every block on its own thread).  general_work()
 * only translates between the runtime's types and the C ABI:
 *   pmt stream tags  <-> c8b_tag   (keys as lib/sync_impl.cc:124-136, lib/signal_impl.cc:135-152, lib/demod_impl.cc:224-263)
 *   message port out <-  PDU records (lib/decode_impl.cc:512-516)
 * which items a call consumes / produces and all arithmetic (CUDA kernels) are decided inside c8b_blk_work().
 * Like the reference's blocks a shell never throws from general_work: errors are printed and the input is dropped.
 */
#include <gnuradio/io_signature.h>
#include <gnuradio/ieee80211/decode.h>
#include <gnuradio/ieee80211/demod.h>
#include <gnuradio/ieee80211/demod2.h>
#include <gnuradio/ieee80211/signal.h>
#include <gnuradio/ieee80211/signal2.h>
#include <gnuradio/ieee80211/sync.h>
#include <gnuradio/ieee80211/trigger.h>

#include <c80211b200.h>

#include <algorithm>
#include <complex>
#include <iostream>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

namespace gr {
namespace ieee80211 {

namespace {

const char* const kNames[7] = { "trigger", "sync", "signal", "signal2", "demod", "demod2", "decode" };

gr::io_signature::sptr port_signature(int kind, bool inputs)
{
    int nin = 0, nout = 0, ib[3] = { 0, 0, 0 }, ob[2] = { 0, 0 };
    c8b_blk_ports(kind, &nin, &nout, ib, ob);
    const int n = inputs ? nin : nout;
    if (n == 0) return gr::io_signature::make(0, 0, 0);
    std::vector<int> sizes(inputs ? ib : ob, (inputs ? ib : ob) + n);
    return gr::io_signature::makev(n, n, sizes);
}

template <class Api, int KIND>
class c8b_block : public Api
{
    c8b_blk* d_blk = nullptr;
    const bool d_debug;
    std::vector<c8b_tag> d_inTags, d_outTags;
    std::vector<uint8_t> d_msg;
    // decode(ifdebug): the counters behind the debug lines (lib/decode_impl.cc:377-411,456-509)
    long d_nPktCorrect = 0, d_legacyMcsCount[8] = { 0 }, d_htMcsCount[8] = { 0 }, d_vhtMcsCount[10] = { 0 };

public:
    c8b_block(int mupos, int mugid, bool debug)
        : gr::block(kNames[KIND], port_signature(KIND, true), port_signature(KIND, false)), d_debug(debug), d_outTags(8), d_msg(1 << 16)
    {
        c8b_cfg cfg = c8b_cfg();
        cfg.mupos = mupos;
        cfg.mugid = mugid;
        if (c8b_blk_create(&cfg, KIND, &d_blk) != C8B_OK)            // no GPU, no block: there is no CPU path
            throw std::runtime_error(std::string("ieee80211 ") + kNames[KIND] + ": " + c8b_blk_last_error(nullptr));
        if (KIND == C8B_BLK_DECODE) this->message_port_register_out(pmt::mp("out"));
        if (KIND >= C8B_BLK_SIGNAL) this->set_tag_propagation_policy(gr::block::TPP_DONT);
    }
    ~c8b_block() override { c8b_blk_destroy(d_blk); }

    void forecast(int noutput_items, gr_vector_int& ninput_items_required) override
    {
        for (auto& r : ninput_items_required) r = c8b_blk_forecast(KIND, noutput_items);
    }

    // decode publishes a frame's MPDUs one call after its last soft bit (the GPU works on it meanwhile): when the flowgraph
    // stops, a call without input collects what is still in flight
    bool stop() override
    {
        if (KIND == C8B_BLK_DECODE) {
            gr_vector_int none(1, 0);
            gr_vector_const_void_star in(1, nullptr);
            gr_vector_void_star out;
            general_work(0, none, in, out);                         // at most four frames in flight: one call publishes them all
        }
        return true;
    }

    int general_work(int noutput_items, gr_vector_int& ninput_items, gr_vector_const_void_star& input_items,
                     gr_vector_void_star& output_items) override
    {
        read_tags(ninput_items[0]);
        int consumed = 0, produced = 0, nTags = 0, nMsg = 0;
        const int rc = c8b_blk_work(d_blk, noutput_items, ninput_items.data(), input_items.data(), output_items.data(), d_inTags.data(),
                                    (int)d_inTags.size(), &consumed, &produced, d_outTags.data(), (int)d_outTags.size(), &nTags,
                                    d_msg.data(), (int)d_msg.size(), &nMsg);
        if (rc != C8B_OK) {
            std::cout << "ieee80211 " << kNames[KIND] << ", error: " << c8b_blk_last_error(d_blk) << std::endl;
            this->consume_each(*std::min_element(ninput_items.begin(), ninput_items.end()));
            return 0;
        }
        for (int k = 0; k < nTags; k++) write_tag(d_outTags[k]);
        for (int o = 0; o + 3 <= nMsg;) {
            const int len = d_msg[o + 1] | (d_msg[o + 2] << 8);
            const int n = d_msg[o] == 20 ? len + 3 : len + 4;         // NDP channel report : PDU record
            pmt::pmt_t meta = pmt::dict_add(pmt::make_dict(), pmt::mp("len"), pmt::from_long(n));
            this->message_port_pub(pmt::mp("out"), pmt::cons(meta, pmt::make_blob(d_msg.data() + o, n)));
            o += n;
        }
        this->consume_each(consumed);
        return produced;
    }

private:
    // tags on input port 0 inside this call's window, grouped by item: one c8b_tag per tagged item
    void read_tags(int ninput0)
    {
        d_inTags.clear();
        if (KIND < C8B_BLK_SIGNAL || ninput0 <= 0) return;
        std::vector<gr::tag_t> tags;
        const uint64_t base = this->nitems_read(0);
        this->get_tags_in_range(tags, 0, base, base + (uint64_t)ninput0);
        std::map<uint64_t, size_t> at;
        for (const gr::tag_t& t : tags) {
            auto it = at.find(t.offset);
            if (it == at.end()) {
                if (d_inTags.size() >= 64) continue;
                c8b_tag z = c8b_tag();
                z.idx = (int32_t)(t.offset - base);
                it = at.emplace(t.offset, d_inTags.size()).first;
                d_inTags.push_back(z);
            }
            set_field(d_inTags[it->second], pmt::symbol_to_string(t.key), t.value);
        }
    }

    static void set_field(c8b_tag& g, const std::string& key, const pmt::pmt_t& v)
    {
        c8b_frame& f = g.f;
        const bool fromSignal = KIND == C8B_BLK_DEMOD || KIND == C8B_BLK_DEMOD2;   // signal's "mcs" / "len" are the L-SIG fields
        if (key == "rad") f.rad = pmt::to_float(v);
        else if (key == "snr") f.snr = pmt::to_float(v);
        else if (key == "rssi") f.rssi = pmt::to_float(v);
        else if (key == "cfo") f.cfo_hz = pmt::to_float(v);
        else if (key == "sssnr0") f.sssnr0 = pmt::to_float(v);
        else if (key == "sssnr1") f.sssnr1 = pmt::to_float(v);
        else if (key == "seq") g.seq = (int32_t)pmt::to_long(v);
        else if (key == "nsamp") f.nsamp = (int32_t)pmt::to_long(v);
        else if (key == "mcs") (fromSignal ? f.l_mcs : f.mcs) = (int32_t)pmt::to_long(v);
        else if (key == "len") (fromSignal ? f.l_len : f.len) = (int32_t)pmt::to_long(v);
        else if (key == "format") f.format = (int32_t)pmt::to_long(v);
        else if (key == "cr") f.cr = (int32_t)pmt::to_long(v);
        else if (key == "ampdu") f.ampdu = (int32_t)pmt::to_long(v);
        else if (key == "trellis") f.trellis = (int32_t)pmt::to_long(v);
        else if (key == "total") f.total = (int32_t)pmt::to_long(v);
        else if (key == "chan" || key == "mu2x1chan") {
            const std::vector<std::complex<float>> c = pmt::c32vector_elements(v);
            g.nvec = (int32_t)std::min<size_t>(c.size(), 128);
            for (int k = 0; k < g.nvec; k++) { g.vec[2 * k] = c[k].real(); g.vec[2 * k + 1] = c[k].imag(); }
        }
    }

    void put(uint64_t at, const char* key, const pmt::pmt_t& v) { this->add_item_tag(0, at, pmt::mp(key), v, this->alias_pmt()); }

    void write_tag(const c8b_tag& g)
    {
        const c8b_frame& f = g.f;
        if (KIND == C8B_BLK_DECODE) { debug_line(f); return; }       // frame report, not a stream tag
        const uint64_t at = this->nitems_written(0) + (uint64_t)g.idx;
        if (KIND == C8B_BLK_SYNC) {                                   // lib/sync_impl.cc:124-136
            put(at, "rad", pmt::from_float(f.rad));
            put(at, "snr", pmt::from_float(f.snr));
            put(at, "rssi", pmt::from_float(f.rssi));
        } else if (KIND == C8B_BLK_SIGNAL || KIND == C8B_BLK_SIGNAL2) {   // lib/signal_impl.cc:135-152
            std::vector<std::complex<float>> h(64);
            for (int k = 0; k < 64; k++) h[k] = std::complex<float>(g.vec[2 * k], g.vec[2 * k + 1]);
            put(at, "cfo", pmt::from_float(f.cfo_hz));
            put(at, "snr", pmt::from_float(f.snr));
            put(at, "rssi", pmt::from_float(f.rssi));
            put(at, "seq", pmt::from_long(g.seq));
            put(at, "mcs", pmt::from_long(f.l_mcs));
            put(at, "len", pmt::from_long(f.l_len));
            put(at, "nsamp", pmt::from_long(f.nsamp));
            put(at, "chan", pmt::init_c32vector(h.size(), h));
        } else {                                                      // lib/demod_impl.cc:224-263, lib/demod2_impl.cc:230-260
            put(at, "cfo", pmt::from_float(f.cfo_hz));
            put(at, "snr", pmt::from_float(f.snr));
            put(at, "rssi", pmt::from_float(f.rssi));
            if (f.format == C8B_F_VHT) {
                put(at, "sssnr0", pmt::from_float(f.sssnr0));
                if (KIND == C8B_BLK_DEMOD2 && f.nss != 1) put(at, "sssnr1", pmt::from_float(f.sssnr1));
            }
            put(at, "format", pmt::from_long(f.format));
            put(at, "mcs", pmt::from_long(f.mcs));
            put(at, "len", pmt::from_long(f.len));
            put(at, "cr", pmt::from_long(f.cr));
            put(at, "ampdu", pmt::from_long(f.ampdu));
            put(at, "trellis", pmt::from_long(f.trellis));
            put(at, "total", pmt::from_long(f.total));
            if (g.nvec == 128) {                                      // NDP: the two VHT-LTFs (lib/demod_impl.cc:238-249)
                std::vector<std::complex<float>> c(128);
                for (int k = 0; k < 128; k++) c[k] = std::complex<float>(g.vec[2 * k], g.vec[2 * k + 1]);
                put(at, "mu2x1chan", pmt::init_c32vector(c.size(), c));
            }
        }
    }

    // decode(ifdebug = true): the lines tools/performance/perf_siso.py:105-118 scrapes
    void debug_line(const c8b_frame& f)
    {
        if (!d_debug) return;
        const bool ok = f.npdu > 0;
        if (!ok && f.format == C8B_F_HT && f.ampdu) return;          // lib/decode_impl.cc:421-424: no line for a failed HT A-MPDU
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

trigger::sptr trigger::make() { return gnuradio::make_block_sptr<c8b_block<trigger, C8B_BLK_TRIGGER>>(0, 0, false); }
sync::sptr sync::make() { return gnuradio::make_block_sptr<c8b_block<sync, C8B_BLK_SYNC>>(0, 0, false); }
signal::sptr signal::make() { return gnuradio::make_block_sptr<c8b_block<signal, C8B_BLK_SIGNAL>>(0, 0, false); }
signal2::sptr signal2::make() { return gnuradio::make_block_sptr<c8b_block<signal2, C8B_BLK_SIGNAL2>>(0, 0, false); }
demod::sptr demod::make(int mupos, int mugid) { return gnuradio::make_block_sptr<c8b_block<demod, C8B_BLK_DEMOD>>(mupos, mugid, false); }
demod2::sptr demod2::make() { return gnuradio::make_block_sptr<c8b_block<demod2, C8B_BLK_DEMOD2>>(0, 0, false); }
decode::sptr decode::make(bool ifdebug) { return gnuradio::make_block_sptr<c8b_block<decode, C8B_BLK_DECODE>>(0, 0, ifdebug); }

}  // namespace ieee80211
}  // namespace gr
