/**
 * kmer_bitset.hpp -- the 128-bit k-mer value type of the drop-in C++ API.
 *
 * The reference declares `typedef boost::dynamic_bitset<> kmer_bitset` (src/kmer.hpp:27) and only ever
 * sizes it to KMER_BITSET_SIZE = 128 bits (src/kmer.hpp:37,53).  Boost is not a dependency here: this
 * class is a trivially copyable two-word value with the dynamic_bitset operations the reference's
 * callers use: size / default constructors, operator[] with a writable proxy,
 * <<= >>= << >> & | ^ ~ |= &= ^=, flip(), count(), ==, != and < (unsigned compare), stream output of
 * size() characters most-significant first, plus std::hash and a Boost-compatible hash_value.
 * Block order matches boost::dynamic_bitset<unsigned long>: word 0 = bits 0..63, word 1 = bits 64..127,
 * which is also how a mask or k-mer crosses the C ABI (include/sks.h).
 */
#ifndef SKS_KMER_BITSET_HPP
#define SKS_KMER_BITSET_HPP

#include <cstddef>
#include <cstdint>
#include <functional>
#include <ostream>

class kmer_bitset
{
public:
    typedef unsigned long block_type;
    static constexpr std::size_t max_bits = 128;

    /** Proxy returned by the non-const operator[] (dynamic_bitset::reference). */
    class reference
    {
        kmer_bitset &b_;
        std::size_t pos_;

    public:
        reference(kmer_bitset &b, std::size_t pos) : b_(b), pos_(pos) {}
        operator bool() const { return b_.test(pos_); }
        bool operator~() const { return !b_.test(pos_); }
        reference &operator=(bool v)
        {
            b_.assign(pos_, v);
            return *this;
        }
        reference &operator=(const reference &o)
        {
            b_.assign(pos_, static_cast<bool>(o));
            return *this;
        }
        reference &operator|=(bool v)
        {
            if (v) b_.assign(pos_, true);
            return *this;
        }
        reference &operator&=(bool v)
        {
            if (!v) b_.assign(pos_, false);
            return *this;
        }
        reference &flip()
        {
            b_.assign(pos_, !b_.test(pos_));
            return *this;
        }
    };

    /** dynamic_bitset's default constructor makes an empty (0-bit) set. */
    kmer_bitset() : w_{0, 0}, nbits_(0) {}
    /** `kmer_bitset(KMER_BITSET_SIZE)`; `value` initialises the low bits like dynamic_bitset does. */
    explicit kmer_bitset(std::size_t num_bits, unsigned long value = 0)
        : w_{value, 0}, nbits_(num_bits > max_bits ? max_bits : static_cast<uint32_t>(num_bits))
    {
        trim();
    }
    /** From the two 64-bit blocks (low word first), as they cross the C ABI. */
    static kmer_bitset from_words(uint64_t lo, uint64_t hi, std::size_t num_bits = max_bits)
    {
        kmer_bitset b(num_bits);
        b.w_[0] = lo;
        b.w_[1] = hi;
        b.trim();
        return b;
    }

    std::size_t size() const { return nbits_; }
    bool empty() const { return nbits_ == 0; }
    uint64_t word(int i) const { return w_[i]; }
    const uint64_t *words() const { return w_; }

    bool test(std::size_t pos) const { return (w_[pos >> 6] >> (pos & 63)) & 1u; }
    bool operator[](std::size_t pos) const { return test(pos); }
    reference operator[](std::size_t pos) { return reference(*this, pos); }
    kmer_bitset &set(std::size_t pos, bool v = true)
    {
        assign(pos, v);
        return *this;
    }
    kmer_bitset &reset()
    {
        w_[0] = w_[1] = 0;
        return *this;
    }

    std::size_t count() const
    {
        return static_cast<std::size_t>(__builtin_popcountll(w_[0]) + __builtin_popcountll(w_[1]));
    }
    bool any() const { return (w_[0] | w_[1]) != 0; }
    bool none() const { return !any(); }

    kmer_bitset &flip()
    {
        w_[0] = ~w_[0];
        w_[1] = ~w_[1];
        trim();
        return *this;
    }
    kmer_bitset operator~() const
    {
        kmer_bitset r(*this);
        r.flip();
        return r;
    }

    kmer_bitset &operator<<=(std::size_t n)
    {
        if (n >= nbits_)
        {
            w_[0] = w_[1] = 0;
        }
        else if (n >= 64)
        {
            w_[1] = w_[0] << (n - 64);
            w_[0] = 0;
        }
        else if (n > 0)
        {
            w_[1] = (w_[1] << n) | (w_[0] >> (64 - n));
            w_[0] <<= n;
        }
        trim();
        return *this;
    }
    kmer_bitset &operator>>=(std::size_t n)
    {
        if (n >= nbits_)
        {
            w_[0] = w_[1] = 0;
        }
        else if (n >= 64)
        {
            w_[0] = w_[1] >> (n - 64);
            w_[1] = 0;
        }
        else if (n > 0)
        {
            w_[0] = (w_[0] >> n) | (w_[1] << (64 - n));
            w_[1] >>= n;
        }
        return *this;
    }
    kmer_bitset operator<<(std::size_t n) const
    {
        kmer_bitset r(*this);
        r <<= n;
        return r;
    }
    kmer_bitset operator>>(std::size_t n) const
    {
        kmer_bitset r(*this);
        r >>= n;
        return r;
    }

    kmer_bitset &operator&=(const kmer_bitset &o)
    {
        w_[0] &= o.w_[0];
        w_[1] &= o.w_[1];
        return *this;
    }
    kmer_bitset &operator|=(const kmer_bitset &o)
    {
        w_[0] |= o.w_[0];
        w_[1] |= o.w_[1];
        trim();
        return *this;
    }
    kmer_bitset &operator^=(const kmer_bitset &o)
    {
        w_[0] ^= o.w_[0];
        w_[1] ^= o.w_[1];
        trim();
        return *this;
    }

    friend kmer_bitset operator&(const kmer_bitset &a, const kmer_bitset &b)
    {
        kmer_bitset r(a);
        r &= b;
        return r;
    }
    friend kmer_bitset operator|(const kmer_bitset &a, const kmer_bitset &b)
    {
        kmer_bitset r(a);
        r |= b;
        return r;
    }
    friend kmer_bitset operator^(const kmer_bitset &a, const kmer_bitset &b)
    {
        kmer_bitset r(a);
        r ^= b;
        return r;
    }
    friend bool operator==(const kmer_bitset &a, const kmer_bitset &b)
    {
        return a.nbits_ == b.nbits_ && a.w_[0] == b.w_[0] && a.w_[1] == b.w_[1];
    }
    friend bool operator!=(const kmer_bitset &a, const kmer_bitset &b) { return !(a == b); }
    /** dynamic_bitset's operator< on equal sizes: comparison as unsigned integers. */
    friend bool operator<(const kmer_bitset &a, const kmer_bitset &b)
    {
        return a.w_[1] != b.w_[1] ? a.w_[1] < b.w_[1] : a.w_[0] < b.w_[0];
    }
    friend bool operator>(const kmer_bitset &a, const kmer_bitset &b) { return b < a; }
    friend bool operator<=(const kmer_bitset &a, const kmer_bitset &b) { return !(b < a); }
    friend bool operator>=(const kmer_bitset &a, const kmer_bitset &b) { return !(a < b); }

    /** size() characters, most significant bit first (src/kmer-sketching.cpp:76 prints masks this way). */
    friend std::ostream &operator<<(std::ostream &os, const kmer_bitset &b)
    {
        for (std::size_t i = b.nbits_; i-- > 0;) os << (b.test(i) ? '1' : '0');
        return os;
    }

private:
    void assign(std::size_t pos, bool v)
    {
        const uint64_t m = uint64_t(1) << (pos & 63);
        if (v)
            w_[pos >> 6] |= m;
        else
            w_[pos >> 6] &= ~m;
    }
    void trim()
    {
        if (nbits_ >= 128) return;
        if (nbits_ >= 64)
        {
            w_[1] &= (nbits_ == 64) ? 0 : ((uint64_t(1) << (nbits_ - 64)) - 1);
        }
        else
        {
            w_[1] = 0;
            w_[0] &= (nbits_ == 0) ? 0 : ((uint64_t(1) << nbits_) - 1);
        }
    }

    uint64_t w_[2];
    uint32_t nbits_;
};

namespace sks
{
/** boost::hash_combine flavours (the reference pins no Boost version; see include/sks.h). */
constexpr int BOOST_HASH_171 = 171; /* Boost 1.71 .. 1.80 */
constexpr int BOOST_HASH_181 = 181; /* Boost >= 1.81, the default */
/** Process-wide selection used by frac_min_hash on the host and by the recognised device predicate. */
int boost_hash_variant();
void set_boost_hash_variant(int variant);
/** hash_value(boost::dynamic_bitset<>) restated: hc(num_bits, hash_range(blocks)). */
std::size_t boost_hash_value(const kmer_bitset &b);
} // namespace sks

namespace std
{
template <>
struct hash<kmer_bitset>
{
    size_t operator()(const kmer_bitset &b) const noexcept
    {
        // only orders the host hash table (src/kmer.hpp:113-124); never result-affecting
        uint64_t x = b.word(0) * 0x9E3779B97F4A7C15ull ^ (b.word(1) + 0xBF58476D1CE4E5B9ull);
        x ^= x >> 29;
        return static_cast<size_t>(x * 0x94D049BB133111EBull) ^ b.size();
    }
};
} // namespace std

#endif // SKS_KMER_BITSET_HPP
