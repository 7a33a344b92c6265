// Internal interface of the verification side of the engine (pairing.cu / verify.cu).
#pragma once
#include "engine.h"

namespace kzg {

// Decompress the 65 G2 points (setup.c:469-477; failure sets *d_bad) and precompute the Miller-loop
// lines of the fixed G2 arguments (G2 generator, [tau]G2, [tau^64]G2).
int setup_g2_and_lines(cudaStream_t stream, Launch& L, Ctx* c, const uint8_t* g2_monomial_bytes_host, int* d_bad);
// is_trusted_setup_in_lagrange_form (setup.c:339-358): pairing check on the first two Lagrange points
// (in file order, i.e. before the bit-reversal).
int setup_is_monomial_form(cudaStream_t stream, Launch& L, Ctx* c, const uint8_t* g1_lagrange_bytes_host, int* is_monomial);

}  // namespace kzg
