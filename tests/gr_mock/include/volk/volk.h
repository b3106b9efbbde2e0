// TEST INFRASTRUCTURE (see pmt/pmt.h): the reference's signal blocks include <volk/volk.h> (lib/signal_impl.h:26,
// lib/signal2_impl.h:26) and call nothing from it.
#pragma once
