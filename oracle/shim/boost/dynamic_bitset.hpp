// TEST INFRASTRUCTURE ONLY -- not part of the product path.
//
// Minimal stand-in for <boost/dynamic_bitset.hpp> so that the UNMODIFIED reference
// sources under /root/reference/src compile in an image that has no Boost.  It provides
// exactly the operator subset the reference uses with Boost's
// semantics: heap (std::vector<Block>) storage like the real class, LSB-first block
// order, unsigned-integer operator<, MSB-first stream output, and
// hash_value(bitset) = hash_combine(hash(num_bits), hash_range(blocks)).
//
// Written from the documented behaviour of Boost.DynamicBitset; no Boost code is copied.
#ifndef ORACLE_SHIM_DYNAMIC_BITSET_HPP
#define ORACLE_SHIM_DYNAMIC_BITSET_HPP

#include <cstddef>
#include <cstdint>
#include <functional>
#include <memory>
#include <ostream>
#include <vector>

#include "functional/hash.hpp"

namespace boost {

template <typename Block = unsigned long, typename Allocator = std::allocator<Block>>
class dynamic_bitset {
 public:
  typedef Block block_type;
  typedef std::size_t size_type;
  static constexpr size_type bits_per_block = sizeof(Block) * 8;

  class reference {
   public:
    reference(Block& b, size_type pos) : blk_(b), mask_(Block(1) << pos) {}
    operator bool() const { return (blk_ & mask_) != 0; }
    bool operator~() const { return (blk_ & mask_) == 0; }
    reference& operator=(bool x) {
      if (x) blk_ |= mask_; else blk_ &= ~mask_;
      return *this;
    }
    reference& operator=(const reference& rhs) { return *this = bool(rhs); }
    reference& flip() { blk_ ^= mask_; return *this; }
   private:
    Block& blk_;
    const Block mask_;
  };

  dynamic_bitset() : nbits_(0) {}
  explicit dynamic_bitset(size_type num_bits, unsigned long value = 0)
      : bits_((num_bits + bits_per_block - 1) / bits_per_block, Block(0)), nbits_(num_bits) {
    if (!bits_.empty()) bits_[0] = Block(value);
    trim();
  }

  size_type size() const { return nbits_; }
  size_type num_blocks() const { return bits_.size(); }

  reference operator[](size_type pos) {
    return reference(bits_[pos / bits_per_block], pos % bits_per_block);
  }
  bool operator[](size_type pos) const {
    return (bits_[pos / bits_per_block] >> (pos % bits_per_block)) & Block(1);
  }
  bool test(size_type pos) const { return (*this)[pos]; }

  dynamic_bitset& operator&=(const dynamic_bitset& r) {
    for (size_type i = 0; i < bits_.size(); ++i) bits_[i] &= r.bits_[i];
    return *this;
  }
  dynamic_bitset& operator|=(const dynamic_bitset& r) {
    for (size_type i = 0; i < bits_.size(); ++i) bits_[i] |= r.bits_[i];
    return *this;
  }
  dynamic_bitset& operator^=(const dynamic_bitset& r) {
    for (size_type i = 0; i < bits_.size(); ++i) bits_[i] ^= r.bits_[i];
    return *this;
  }

  dynamic_bitset& operator<<=(size_type n) {
    if (n >= nbits_) { for (auto& b : bits_) b = 0; return *this; }
    if (n == 0) return *this;
    const size_type nb = bits_.size(), div = n / bits_per_block, r = n % bits_per_block;
    for (size_type i = nb; i-- > 0;) {
      Block v = 0;
      if (i >= div) {
        v = bits_[i - div] << r;
        if (r && i >= div + 1) v |= bits_[i - div - 1] >> (bits_per_block - r);
      }
      bits_[i] = v;
    }
    trim();
    return *this;
  }
  dynamic_bitset& operator>>=(size_type n) {
    if (n >= nbits_) { for (auto& b : bits_) b = 0; return *this; }
    if (n == 0) return *this;
    const size_type nb = bits_.size(), div = n / bits_per_block, r = n % bits_per_block;
    for (size_type i = 0; i < nb; ++i) {
      Block v = 0;
      if (i + div < nb) {
        v = bits_[i + div] >> r;
        if (r && i + div + 1 < nb) v |= bits_[i + div + 1] << (bits_per_block - r);
      }
      bits_[i] = v;
    }
    return *this;
  }
  dynamic_bitset operator<<(size_type n) const { dynamic_bitset t(*this); return t <<= n; }
  dynamic_bitset operator>>(size_type n) const { dynamic_bitset t(*this); return t >>= n; }

  dynamic_bitset& flip() {
    for (auto& b : bits_) b = ~b;
    trim();
    return *this;
  }
  dynamic_bitset operator~() const { dynamic_bitset t(*this); t.flip(); return t; }

  size_type count() const {
    size_type c = 0;
    for (Block b : bits_) c += static_cast<size_type>(__builtin_popcountll(b));
    return c;
  }
  bool any() const { for (Block b : bits_) if (b) return true; return false; }
  bool none() const { return !any(); }

  friend bool operator==(const dynamic_bitset& a, const dynamic_bitset& b) {
    return a.nbits_ == b.nbits_ && a.bits_ == b.bits_;
  }
  friend bool operator!=(const dynamic_bitset& a, const dynamic_bitset& b) { return !(a == b); }
  // Equal sizes: compare as unsigned integers, most significant block first.
  friend bool operator<(const dynamic_bitset& a, const dynamic_bitset& b) {
    for (size_type i = a.bits_.size(); i-- > 0;) {
      if (a.bits_[i] != b.bits_[i]) return a.bits_[i] < b.bits_[i];
    }
    return false;
  }
  friend bool operator>(const dynamic_bitset& a, const dynamic_bitset& b) { return b < a; }
  friend bool operator<=(const dynamic_bitset& a, const dynamic_bitset& b) { return !(b < a); }
  friend bool operator>=(const dynamic_bitset& a, const dynamic_bitset& b) { return !(a < b); }

  friend dynamic_bitset operator&(const dynamic_bitset& a, const dynamic_bitset& b) {
    dynamic_bitset t(a); return t &= b;
  }
  friend dynamic_bitset operator|(const dynamic_bitset& a, const dynamic_bitset& b) {
    dynamic_bitset t(a); return t |= b;
  }
  friend dynamic_bitset operator^(const dynamic_bitset& a, const dynamic_bitset& b) {
    dynamic_bitset t(a); return t ^= b;
  }

  friend std::ostream& operator<<(std::ostream& os, const dynamic_bitset& b) {
    for (size_type i = b.nbits_; i-- > 0;) os << (b[i] ? '1' : '0');
    return os;
  }

  // Boost >= 1.71: hash_value(dynamic_bitset) = combine(hash(m_num_bits), hash(m_bits)).
  friend std::size_t hash_value(const dynamic_bitset& a) {
    std::size_t res = boost::hash_value(a.nbits_);
    boost::hash_combine(res, a.bits_);
    return res;
  }

  // Direct block access for the test glue (to_block_range equivalent).
  const std::vector<Block, Allocator>& shim_blocks() const { return bits_; }
  std::vector<Block, Allocator>& shim_blocks() { return bits_; }

 private:
  void trim() {
    const size_type extra = nbits_ % bits_per_block;
    if (extra && !bits_.empty()) bits_.back() &= (Block(1) << extra) - 1;
  }
  std::vector<Block, Allocator> bits_;
  size_type nbits_;
};

}  // namespace boost

namespace std {
template <typename B, typename A>
struct hash<boost::dynamic_bitset<B, A>> {
  size_t operator()(const boost::dynamic_bitset<B, A>& a) const { return hash_value(a); }
};
}  // namespace std

#endif
