// TEST INFRASTRUCTURE (see pmt/pmt.h)
#pragma once
#define __GR_ATTR_EXPORT __attribute__((visibility("default")))
#define __GR_ATTR_IMPORT __attribute__((visibility("default")))
