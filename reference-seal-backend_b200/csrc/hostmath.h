// hostmath.h -- host-side number theory used to build the device tables (twiddles, Barrett /
// Shoup constants, RNS base-conversion matrices).  Runs once per context on the CPU; none of it
// is ciphertext arithmetic.
#pragma once
#include <stdint.h>

#include <vector>

namespace b200he {
namespace hm {

typedef uint64_t u64;
typedef unsigned __int128 u128;

inline u64 mulmod(u64 a, u64 b, u64 q) { return (u64)((u128)a * b % q); }
inline u64 powmod(u64 a, u64 e, u64 q)
{
    u64 r = 1 % q;
    for (a %= q; e; e >>= 1, a = mulmod(a, a, q))
        if (e & 1) r = mulmod(r, a, q);
    return r;
}
inline u64 invmod(u64 a, u64 q) { return powmod(a % q, q - 2, q); }   // q prime
inline u64 shoup(u64 w, u64 q) { return (u64)(((u128)w << 64) / q); }
inline int bitlen(u64 v)
{
    int b = 0;
    for (; v; v >>= 1) b++;
    return b;
}
inline uint32_t brv(uint32_t x, int bits)
{
    uint32_t r = 0;
    for (int i = 0; i < bits; i++, x >>= 1) r = (r << 1) | (x & 1u);
    return r;
}
inline bool is_prime(u64 n)
{
    if (n < 2) return false;
    static const u64 bases[] = { 2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37 };
    for (u64 p : bases) {
        if (n == p) return true;
        if (n % p == 0) return false;
    }
    u64 d = n - 1;
    int r = 0;
    while (!(d & 1)) d >>= 1, r++;
    for (u64 a : bases) {
        u64 x = powmod(a, d, n);
        if (x == 1 || x == n - 1) continue;
        bool composite = true;
        for (int i = 1; i < r && composite; i++) {
            x = mulmod(x, x, n);
            if (x == n - 1) composite = false;
        }
        if (composite) return false;
    }
    return true;
}
// descending primes p = 1 (mod factor), 2^(bits-1) < p < 2^bits
inline std::vector<u64> primes_below(u64 factor, int bits, size_t count)
{
    std::vector<u64> out;
    u64 v = ((u64(1) << bits) - 1) / factor * factor + 1, lo = u64(1) << (bits - 1);
    for (; out.size() < count && v > lo; v -= factor)
        if (is_prime(v)) out.push_back(v);
    return out;
}
// some primitive 2N-th root of unity mod q (used for the BEHZ auxiliary primes, whose transforms
// are internal: any primitive root yields the same products)
inline u64 any_primitive_root(u64 two_n, u64 q)
{
    if ((q - 1) % two_n) return 0;
    u64 e = (q - 1) / two_n;
    for (u64 g = 2; g < 100000; g++) {
        u64 x = powmod(g, e, q);
        if (powmod(x, two_n / 2, q) == q - 1) return x;
    }
    return 0;
}
// bit length of the product of the given factors
inline size_t product_bits(const std::vector<u64> &f)
{
    std::vector<u64> w(1, 1);
    for (u64 x : f) {
        u64 carry = 0;
        for (auto &limb : w) {
            u128 p = (u128)limb * x + carry;
            limb = (u64)p;
            carry = (u64)(p >> 64);
        }
        if (carry) w.push_back(carry);
    }
    return (w.size() - 1) * 64 + (size_t)bitlen(w.back());
}

}   // namespace hm
}   // namespace b200he
