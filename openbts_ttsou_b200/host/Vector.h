// host/Vector.h -- fixed-size owning/aliasing array, the subset of the reference's CommonLibs/Vector.h semantics
// the sigProcLib surface relies on: size(), begin()/end(), operator[], fill(), segment aliases that do not own,
// copy construction that clones, segmentCopyTo/copyToSegment.  Written for stand-alone builds of the shim.
#ifndef BTSDSP_HOST_VECTOR_H
#define BTSDSP_HOST_VECTOR_H
#include <stddef.h>
#include <string.h>

template <class T> class Vector {
 protected:
  T *mData;    // owned block or NULL for aliases
  T *mStart, *mEnd;

 public:
  typedef T *iterator;
  typedef const T *const_iterator;

  explicit Vector(size_t n = 0) : mData(NULL), mStart(NULL), mEnd(NULL) { resize(n); }
  Vector(const Vector<T> &o) : mData(NULL), mStart(NULL), mEnd(NULL) { clone(o); }
  Vector(T *data, T *start, T *end) : mData(data), mStart(start), mEnd(end) {}          // explicit (maybe alias)
  Vector(T *start, size_t span) : mData(NULL), mStart(start), mEnd(start + span) {}     // alias, not owned
  Vector(const Vector<T> &a, const Vector<T> &b) : mData(NULL), mStart(NULL), mEnd(NULL) {
    resize(a.size() + b.size());
    memcpy(mStart, a.mStart, a.bytes());
    memcpy(mStart + a.size(), b.mStart, b.bytes());
  }
  ~Vector() { delete[] mData; }
  Vector<T> &operator=(const Vector<T> &o) { if (this != &o) clone(o); return *this; }

  void resize(size_t n) {
    delete[] mData;
    mData = n ? new T[n]() : NULL;
    mStart = mData;
    mEnd = mStart + n;
  }
  void clone(const Vector<T> &o) { resize(o.size()); if (o.size()) memcpy(mStart, o.mStart, o.bytes()); }
  size_t size() const { return mEnd - mStart; }
  size_t bytes() const { return size() * sizeof(T); }
  T *begin() { return mStart; }
  T *end() { return mEnd; }
  const T *begin() const { return mStart; }
  const T *end() const { return mEnd; }
  T &operator[](size_t k) { return mStart[k]; }
  const T &operator[](size_t k) const { return mStart[k]; }
  void fill(const T &v) { for (T *p = mStart; p < mEnd; ++p) *p = v; }
  Vector<T> segment(size_t start, size_t span) { return Vector<T>(NULL, mStart + start, mStart + start + span); }
  const Vector<T> segment(size_t start, size_t span) const { return Vector<T>(NULL, mStart + start, mStart + start + span); }
  void copyTo(Vector<T> &o) const { memcpy(o.mStart, mStart, bytes()); }
  void copyToSegment(Vector<T> &o, size_t start, size_t span) const { memcpy(o.mStart + start, mStart, span * sizeof(T)); }
  void segmentCopyTo(Vector<T> &o, size_t start, size_t span) const { memcpy(o.mStart, mStart + start, span * sizeof(T)); }
};
#endif
