// Device QMC sequence generation (qmc.cu) -- interface to api.cu.
#pragma once

#include <cstdint>
#include <vector>
#include <cuda_runtime.h>

namespace rtm
{

constexpr int kQmcPrimes = 1000; // PRIME_TBL_SIZE (reference sampling.h:18)

// values of cuda_trace_qmc_sequence's `kind` / `scramble` (include/cuda_trace.h)
enum { kQmcHalton = 0, kQmcHammersley = 1, kQmcHaltonZaremba = 2, kQmcHammersleyZaremba = 3, kQmcBase2 = 4,
       kQmcSobol = 5, kQmcLarcherPillichshammer = 6 };
enum { kQmcScrambleNone = 0, kQmcScrambleBraatenWeller = 1, kQmcScrambleFaure = 2, kQmcScrambleReverse = 3,
       kQmcScrambleCustom = 4 };

std::vector<uint32_t> qmc_primes();
std::vector<uint32_t> qmc_faure_permutation(uint32_t base);
cudaError_t qmc_upload_primes();
void launch_qmc_sequence(uint32_t kind, uint32_t scramble, uint32_t n_begin, uint32_t count, uint32_t dim_begin,
                         uint32_t dim_count, uint32_t num_smp, uint32_t bits, const uint32_t *d_perm,
                         const uint32_t *d_perm_offset, uint32_t perm_primes, double *d_out, cudaStream_t stream);
void launch_cranley_patterson(const double *d_x, double e, uint32_t n, double *d_out, cudaStream_t stream);

} // namespace rtm
