// Low-discrepancy sequences on the device: the rest of the reference's sampling module
// (sampling.cpp:55-290, sampling.h:91-120).  K1 itself only needs the
// unscrambled Hammersley pair (K2 in trace_kernels.cu); this file generates whole tables of
//   Halton / Hammersley with none | Braaten-Weller | Faure | reverse | caller-supplied digit
//   permutations, their folded (Zaremba) forms, and the base-2 radical inverse, Sobol and
//   Larcher-Pillichshammer sequences with XOR scrambling,
// one thread per (sample, dimension), in fp64 with the reference's operation order (this file is
// compiled -fmad=false like the rest, so a*b+c never fuses).
#include "qmc.cuh"

#include <vector>

namespace rtm
{

namespace
{

__constant__ uint32_t c_prime[kQmcPrimes];

// sampling.cpp:194-210.  perm == nullptr: plain digits
__device__ double radical_inverse(uint32_t n, uint32_t base, const uint32_t *__restrict__ perm)
{
    const double inv_base = 1.0 / (double) base;
    double inv_base_i = inv_base, val = 0.0;
    while (n > 0)
    {
        uint32_t digit = n % base;
        if (perm)
            digit = perm[digit];
        val += digit * inv_base_i;
        inv_base_i *= inv_base;
        n /= base;
    }
    return val;
}

// sampling.cpp:251-269: the loop runs until adding the next digit weight no longer changes n
__device__ double folded_radical_inverse(uint32_t n, uint32_t base)
{
    const double inv_base = 1.0 / (double) base;
    double inv_base_i = inv_base, val = 0.0;
    uint32_t offset = 0;
    while ((double) n + base * inv_base_i != (double) n)
    {
        const uint32_t digit = (n + offset) % base;
        val += digit * inv_base_i;
        inv_base_i *= inv_base;
        n /= base;
        offset++;
    }
    return val;
}

__device__ double two_pow_minus32(uint32_t bits) { return (double) bits / 4294967296.0; }

struct QmcParams
{
    uint32_t kind, scramble;
    uint32_t n_begin, count, dim_begin, dim_count, num_smp, bits;
    const uint32_t *perm;        // concatenated permutation tables of the first perm_primes primes
    const uint32_t *perm_offset; // perm_primes + 1 offsets
    uint32_t perm_primes;
    double *out;
};

__global__ void qmc_sequence_kernel(const QmcParams p)
{
    const uint64_t idx = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (uint64_t) p.count * p.dim_count)
        return;
    const uint32_t n = p.n_begin + (uint32_t) (idx / p.dim_count), dim = p.dim_begin + (uint32_t) (idx % p.dim_count);
    double v = 0.0;
    switch (p.kind)
    {
        case kQmcHalton:
        case kQmcHammersley:
        {
            if (p.kind == kQmcHammersley && dim == 0)
            {
                v = (double) n / (double) p.num_smp; // sampling.h:116-117
                break;
            }
            // sampling.h:107-120: Halton uses prime[dim]; Hammersley uses prime[dim - 1] but -- as in
            // the reference -- still the permutation table of index `dim`
            const uint32_t base = c_prime[p.kind == kQmcHalton ? dim : dim - 1];
            const bool has_table = p.scramble != kQmcScrambleNone && dim < p.perm_primes;
            const uint32_t *perm = (has_table && p.scramble != kQmcScrambleReverse) ? p.perm + p.perm_offset[dim] : nullptr;
            // the reverse "table" of index dim is base' - digit with base' = prime[dim] (sampling.cpp:160-171)
            if (has_table && p.scramble == kQmcScrambleReverse)
            {
                const uint32_t table_base = c_prime[dim];
                const double inv_base = 1.0 / (double) base;
                double inv_base_i = inv_base;
                uint32_t m = n;
                while (m > 0)
                {
                    uint32_t digit = m % base;
                    digit = digit ? table_base - digit : 0u;
                    v += digit * inv_base_i;
                    inv_base_i *= inv_base;
                    m /= base;
                }
            }
            else
                v = radical_inverse(n, base, perm);
            break;
        }
        case kQmcHaltonZaremba:
            v = folded_radical_inverse(n, c_prime[dim]); // sampling.cpp:271-274
            break;
        case kQmcHammersleyZaremba:
            // sampling.cpp:276-281: dimension 0 divides in FLOAT
            v = dim == 0 ? (double) ((float) n / (float) p.num_smp) : folded_radical_inverse(n, c_prime[dim - 1]);
            break;
        case kQmcBase2:
            v = two_pow_minus32(__brev(n) ^ p.bits); // sampling.cpp:212-229: bit reversal, XOR scramble
            break;
        case kQmcSobol:
        {
            uint32_t s = p.bits, m = n;
            for (uint32_t w = 1u << 31; m != 0; m >>= 1, w ^= w >> 1) // sampling.cpp:231-239
                if (m & 1u) s ^= w;
            v = two_pow_minus32(s);
            break;
        }
        default: // kQmcLarcherPillichshammer, sampling.cpp:241-249
        {
            uint32_t s = p.bits, m = n;
            for (uint32_t w = 1u << 31; m != 0; m >>= 1, w |= w >> 1)
                if (m & 1u) s ^= w;
            v = two_pow_minus32(s);
            break;
        }
    }
    p.out[idx] = v;
}

// sampling.cpp:283-290, including its test of x (not of x + e) against 1
__global__ void cranley_patterson_kernel(const double *__restrict__ x, double e, uint32_t n, double *__restrict__ out)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n)
        return;
    const double r = x[i] + e;
    out[i] = x[i] > 1.0 ? r - 1.0 : r;
}

} // namespace

// first kQmcPrimes primes by a sieve (the reference ships them as a table, primes.cpp)
std::vector<uint32_t> qmc_primes()
{
    std::vector<uint32_t> primes;
    const uint32_t limit = 8000; // the 1000th prime is 7919
    std::vector<bool> composite(limit + 1, false);
    for (uint32_t i = 2; i <= limit && primes.size() < (size_t) kQmcPrimes; i++)
    {
        if (composite[i])
            continue;
        primes.push_back(i);
        for (uint32_t j = i * i; j <= limit; j += i)
            composite[j] = true;
    }
    return primes;
}

// Faure permutation of one base (Keller, "Monte Carlo and Beyond"; the reference's recursion is
// sampling.cpp:100-133): sigma_2 = identity; even b: (2 sigma_{b/2}, 2 sigma_{b/2} + 1); odd b:
// sigma_{b-1} with the values >= (b-1)/2 incremented and (b-1)/2 inserted in the middle.
std::vector<uint32_t> qmc_faure_permutation(uint32_t base)
{
    if (base == 2)
        return { 0u, 1u };
    std::vector<uint32_t> out;
    if (base % 2 == 0)
    {
        const std::vector<uint32_t> half = qmc_faure_permutation(base / 2);
        for (uint32_t v : half) out.push_back(2 * v);
        for (uint32_t v : half) out.push_back(2 * v + 1);
        return out;
    }
    const std::vector<uint32_t> prev = qmc_faure_permutation(base - 1);
    const uint32_t mid = (base - 1) / 2;
    for (uint32_t i = 0; i < base; i++)
    {
        if (i == mid)
            out.push_back(mid);
        else
        {
            const uint32_t v = prev[i < mid ? i : i - 1];
            out.push_back(v >= mid ? v + 1 : v);
        }
    }
    return out;
}

cudaError_t qmc_upload_primes()
{
    const std::vector<uint32_t> primes = qmc_primes();
    return cudaMemcpyToSymbol(c_prime, primes.data(), sizeof(uint32_t) * kQmcPrimes);
}

void launch_qmc_sequence(uint32_t kind, uint32_t scramble, uint32_t n_begin, uint32_t count, uint32_t dim_begin,
                         uint32_t dim_count, uint32_t num_smp, uint32_t bits, const uint32_t *d_perm,
                         const uint32_t *d_perm_offset, uint32_t perm_primes, double *d_out, cudaStream_t stream)
{
    QmcParams p;
    p.kind = kind; p.scramble = scramble; p.n_begin = n_begin; p.count = count; p.dim_begin = dim_begin;
    p.dim_count = dim_count; p.num_smp = num_smp; p.bits = bits; p.perm = d_perm; p.perm_offset = d_perm_offset;
    p.perm_primes = perm_primes; p.out = d_out;
    const uint64_t total = (uint64_t) count * dim_count;
    if (total)
        qmc_sequence_kernel<<<(unsigned) ((total + 255) / 256), 256, 0, stream>>>(p);
}

void launch_cranley_patterson(const double *d_x, double e, uint32_t n, double *d_out, cudaStream_t stream)
{
    if (n)
        cranley_patterson_kernel<<<(n + 255) / 256, 256, 0, stream>>>(d_x, e, n, d_out);
}

} // namespace rtm
