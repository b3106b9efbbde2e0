// TEST INFRASTRUCTURE: stand-in for the reference's include/gnuradio/ieee80211/api.h (export macro), used when
// /root/reference is not mounted.
#pragma once
#include <gnuradio/attributes.h>
#ifdef gnuradio_ieee80211_EXPORTS
#define IEEE80211_API __GR_ATTR_EXPORT
#else
#define IEEE80211_API __GR_ATTR_IMPORT
#endif
