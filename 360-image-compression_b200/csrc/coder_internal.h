// Packed CDF rows exchanged between the fused codec kernels (codec.cu) and the host coder (coder.cpp).
// The tables are the same integers the per-op path produces as float rows; packing only drops the constant end
// points T[0] = 0 and T[last] = 65536 and stores the rest as 16-bit words plus one overflow bit each (a bin can
// reach 65536 + a few counts before the monotonic fix-up pulls it back, entropy_gmm_table_cuda.cu:85-107).
//   code stream row  :  8 x u16 = T[1..7] low words, meta = sym(3 bits) | mask << 8 | overflow(T[1..7]) << 9
//   importance row   : 64 x u16 = T[1..48] low words, [48] = sym, [49..51] = overflow bits of T[1..48]
#pragma once
#include <stdint.h>
#include "../../include/lic360_b200.h"

namespace lic360 {
int coder_encode_packed_gmm(lic360_coder* c, const uint16_t* rows, int nrows);
int coder_decode_packed_gmm(lic360_coder* c, const uint16_t* rows, int nrows, float* out);
int coder_encode_packed_imp(lic360_coder* c, const uint16_t* rows, int nrows);
int coder_decode_packed_imp(lic360_coder* c, const uint16_t* rows, int nrows, float* out);
const uint8_t* coder_bytes(lic360_coder* c, long* n);
}  // namespace lic360
