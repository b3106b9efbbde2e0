// TEST INFRASTRUCTURE: stand-ins for the reference's seven public block headers (class name, virtual gr::block base and
// make() signature -- all that lib/rx_blocks_impl.cc needs from them).  Only used when /root/reference is not mounted;
// with it, the shells are compiled against the reference's own include/gnuradio/ieee80211/*.h.
#pragma once
#include <gnuradio/block.h>

#define C8B_STANDIN_BLOCK(NAME, ...)                           \
    class NAME : virtual public gr::block                      \
    {                                                          \
    public:                                                    \
        typedef std::shared_ptr<NAME> sptr;                    \
        static sptr make(__VA_ARGS__);                         \
    }

namespace gr {
namespace ieee80211 {
C8B_STANDIN_BLOCK(trigger);
C8B_STANDIN_BLOCK(sync);
C8B_STANDIN_BLOCK(signal);
C8B_STANDIN_BLOCK(signal2);
C8B_STANDIN_BLOCK(demod, int mupos, int mugid);
C8B_STANDIN_BLOCK(demod2);
C8B_STANDIN_BLOCK(decode, bool ifdebug);
}  // namespace ieee80211
}  // namespace gr
