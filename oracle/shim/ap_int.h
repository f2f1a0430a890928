// TEST INFRASTRUCTURE (oracle) - never used by the product path.
//
// Minimal stand-in for the Xilinx `ap_uint<W>` as the reference's HLS kernel
// uses it: burst words of `burst width` bits that are sliced into elements and
// assembled from them with the range operator `word(msb, lsb)`
// (reference: src/soda/codegen/xilinx/hls_kernel.py:444-491,
// `ap_uint<BW>* bank_<b>_<name>` ports at :62-66).  Only what that kernel text
// needs: construction from integers, range read / write of up to 64 bits,
// conversion between widths.  Storage is exactly W / 8 bytes, least
// significant byte first, so that an `ap_uint<BW>*` port can point at the raw
// bank buffers the host tiler wrote and an `ap_uint<32>` has the bytes of a
// `float`.
#pragma once

#include <cstdint>
#include <cstring>
#include <type_traits>

template <int W>
struct ap_uint {
  static_assert(W >= 1, "width");
  static constexpr int kBytes = W <= 8 ? 1 : W <= 16 ? 2 : W <= 32 ? 4
                                : (W + 63) / 64 * 8;
  unsigned char bytes[kBytes];

  ap_uint() { std::memset(bytes, 0, kBytes); }
  ap_uint(uint64_t value) {  // NOLINT: implicit like the original
    std::memset(bytes, 0, kBytes);
    set_bits(0, W < 64 ? W : 64, value);
  }
  template <int V>
  ap_uint(const ap_uint<V>& other) {  // NOLINT
    std::memset(bytes, 0, kBytes);
    constexpr int kCopy = (V < W ? V : W);
    static_assert(kCopy <= 64 || V == W, "wide conversions are not needed");
    if (V == W) {
      std::memcpy(bytes, other.bytes, kBytes < ap_uint<V>::kBytes
                                          ? kBytes : ap_uint<V>::kBytes);
    } else {
      set_bits(0, kCopy, other.get_bits(0, kCopy));
    }
  }

  // bits [lsb, lsb + width), width <= 64
  uint64_t get_bits(int lsb, int width) const {
    uint64_t value = 0;
    if (lsb % 8 == 0 && width % 8 == 0) {
      std::memcpy(&value, bytes + lsb / 8, width / 8);
      return value;
    }
    for (int i = 0; i < width; ++i) {
      const int bit = lsb + i;
      value |= uint64_t((bytes[bit / 8] >> (bit % 8)) & 1u) << i;
    }
    return value;
  }
  void set_bits(int lsb, int width, uint64_t value) {
    if (lsb % 8 == 0 && width % 8 == 0) {
      std::memcpy(bytes + lsb / 8, &value, width / 8);
      return;
    }
    for (int i = 0; i < width; ++i) {
      const int bit = lsb + i;
      const unsigned char mask = static_cast<unsigned char>(1u << (bit % 8));
      if ((value >> i) & 1u) {
        bytes[bit / 8] |= mask;
      } else {
        bytes[bit / 8] &= static_cast<unsigned char>(~mask);
      }
    }
  }

  operator uint64_t() const { return get_bits(0, W < 64 ? W : 64); }  // NOLINT

  struct Range {
    ap_uint& word;
    int msb, lsb;
    template <int V>
    operator ap_uint<V>() const {  // NOLINT
      ap_uint<V> out;
      out.set_bits(0, V, word.get_bits(lsb, msb - lsb + 1));
      return out;
    }
    operator uint64_t() const { return word.get_bits(lsb, msb - lsb + 1); }  // NOLINT
    template <int V>
    Range& operator=(const ap_uint<V>& value) {
      word.set_bits(lsb, msb - lsb + 1, value.get_bits(0, V < 64 ? V : 64));
      return *this;
    }
    Range& operator=(uint64_t value) {
      word.set_bits(lsb, msb - lsb + 1, value);
      return *this;
    }
  };
  struct ConstRange {
    const ap_uint& word;
    int msb, lsb;
    template <int V>
    operator ap_uint<V>() const {  // NOLINT
      ap_uint<V> out;
      out.set_bits(0, V, word.get_bits(lsb, msb - lsb + 1));
      return out;
    }
    operator uint64_t() const { return word.get_bits(lsb, msb - lsb + 1); }  // NOLINT
  };
  Range operator()(int msb, int lsb) { return Range{*this, msb, lsb}; }
  ConstRange operator()(int msb, int lsb) const {
    return ConstRange{*this, msb, lsb};
  }
};
