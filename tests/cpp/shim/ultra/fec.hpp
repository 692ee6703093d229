// tests/cpp/shim/ultra/fec.hpp — TEST INFRASTRUCTURE.  Shadows the reference's include/ultra/fec.hpp when the reference's own test
// sources are compiled UNMODIFIED against the drop-in classes (oracle/ref_build/Makefile: test_multiblock_ldpc_pu): the names the
// reference's tests use resolve to the pu:: classes of include/pu/pu_dropin.hpp, i.e. to libpu_b200.so.
#pragma once
#ifndef PU_DROPIN_WITH_ULTRA
#define PU_DROPIN_WITH_ULTRA
#endif
#include "pu/pu_dropin.hpp"

namespace ultra {
using LDPCEncoder = pu::LDPCEncoder;
using LDPCDecoder = pu::LDPCDecoder;
using Interleaver = pu::Interleaver;
using ChannelInterleaver = pu::ChannelInterleaver;
}  // namespace ultra
