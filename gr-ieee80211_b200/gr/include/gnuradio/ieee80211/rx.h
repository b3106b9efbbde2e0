/* rx.h -- gr::ieee80211::rx: the whole 20 MHz receive chain  presiso -> trigger -> sync -> signal[2] -> demod[2] -> decode
 * of examples/rx.grc:753-767 (nant = 1) / examples/rx2.grc:676-692 (nant = 2) as ONE sink block on the B200.
 *
 * A new block next to the reference's seven (which stay available, lib/rx_blocks_impl.cc): same input (the capture as
 * gr_complex, one stream per antenna), same message port "out" with the same PDU messages as ieee80211::decode
 * (lib/decode_impl.cc:512-516), so network.socket_pdu / tools/macExampleGrRx.py behind it keep working.  It forwards what the
 * scheduler hands it to the library's live-stream session (c8b_stream_push); the device keeps the window and the state the
 * seven blocks keep in their members.  This is the throughput deployment: one host->device copy per work() call instead of
 * a round trip per block.
 */
#ifndef INCLUDED_IEEE80211_RX_H
#define INCLUDED_IEEE80211_RX_H

#include <gnuradio/block.h>
#include <gnuradio/ieee80211/api.h>

namespace gr {
namespace ieee80211 {

class IEEE80211_API rx : virtual public gr::block
{
public:
    typedef std::shared_ptr<rx> sptr;
    /* nant: 1 or 2 receive antennas; mupos / mugid as demod::make(mupos, mugid); ifdebug as decode::make(ifdebug) */
    static sptr make(int nant = 1, int mupos = 0, int mugid = 2, bool ifdebug = false);
};

}  // namespace ieee80211
}  // namespace gr

#endif
