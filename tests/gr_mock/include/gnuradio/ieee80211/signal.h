// TEST INFRASTRUCTURE: stand-in for the reference's public header include/gnuradio/ieee80211/signal.h (class name, base and
// make() signature only), used when /root/reference is not there to compile the shells against the real one.
#pragma once
#include <gnuradio/block.h>
namespace gr { namespace ieee80211 {
class signal : virtual public gr::block { public: typedef std::shared_ptr<signal> sptr; static sptr make(); };
} }
