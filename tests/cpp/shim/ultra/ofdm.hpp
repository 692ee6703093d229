// tests/cpp/shim/ultra/ofdm.hpp — TEST INFRASTRUCTURE: see fec.hpp in this directory.
#pragma once
#ifndef PU_DROPIN_WITH_ULTRA
#define PU_DROPIN_WITH_ULTRA
#endif
#include "pu/pu_dropin.hpp"

namespace ultra {
using OFDMModulator = pu::OFDMModulator;
using OFDMDemodulator = pu::OFDMDemodulator;
}  // namespace ultra
