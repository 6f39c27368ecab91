// host/BitVector.h -- BitVector (one char per bit, value in bit 0) and SoftVector (one float per bit in [0,1],
// hard decision = value > 0.5F), the two container types on the sigProcLib.h boundary
// (reference CommonLibs/BitVector.h:228-337, :352-425).  Only what the burst-DSP surface uses.
#ifndef BTSDSP_HOST_BITVECTOR_H
#define BTSDSP_HOST_BITVECTOR_H
#include "Vector.h"

class BitVector : public Vector<char> {
 public:
  explicit BitVector(size_t n = 0) : Vector<char>(n) {}
  BitVector(const Vector<char> &v) : Vector<char>(v) {}
  BitVector(char *data, char *start, char *end) : Vector<char>(data, start, end) {}
  BitVector(const char *zerosAndOnes) : Vector<char>(strlen(zerosAndOnes)) {
    for (size_t k = 0; k < size(); k++) mStart[k] = (zerosAndOnes[k] == '1');
  }
  bool bit(size_t k) const { return mStart[k] & 0x01; }
  BitVector segment(size_t start, size_t span) { return BitVector(NULL, mStart + start, mStart + start + span); }
  const BitVector segment(size_t start, size_t span) const {
    return BitVector(NULL, const_cast<char *>(mStart) + start, const_cast<char *>(mStart) + start + span);
  }
};

class SoftVector : public Vector<float> {
 public:
  explicit SoftVector(size_t n = 0) : Vector<float>(n) {}
  SoftVector(const Vector<float> &v) : Vector<float>(v) {}
  bool bit(size_t k) const { return mStart[k] > 0.5F; }
  BitVector sliced() const {
    BitVector b(size());
    for (size_t k = 0; k < size(); k++) b[k] = bit(k);
    return b;
  }
};
#endif
