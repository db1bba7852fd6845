// drops_b200/hostxx/misc/container.h -- the one container of DROPS' misc/container.h that the
// Lanczos/MINRES interface exposes publicly: the fixed-size ring buffer SBufferCL
// (reference: src/misc/container.h:201-251).  Index k is taken modulo the size relative to a moving
// front; rotate() makes the old element [1] the new [0] (and the old [0] the new [-1]).
#ifndef DROPS_CONTAINER_H
#define DROPS_CONTAINER_H

#include <algorithm>

#include "misc/utils.h"

namespace DROPS
{

template <class T, Uint _Size>
class SBufferCL
{
  private:
    T slot_[_Size];
    int front_;

    int wrap(int i) const
    {
        if (i < 0) i += static_cast<int>(_Size);
        else if (i >= static_cast<int>(_Size)) i -= static_cast<int>(_Size);
        i += front_;
        return i < static_cast<int>(_Size) ? i : i - static_cast<int>(_Size);
    }

  public:
    typedef T value_type;
    typedef T& reference;

    explicit SBufferCL(T val = T()) : front_(0) { std::fill_n(slot_ + 0, _Size, val); }

    T&       operator[](int i)       { return slot_[wrap(i)]; }
    const T& operator[](int i) const { return slot_[wrap(i)]; }
    Uint size() const { return _Size; }

    /// Overwrites the front element and advances.
    void push_back(const T& t)
    {
        slot_[front_] = t;
        if (++front_ == static_cast<int>(_Size)) front_ = 0;
    }
    /// For positive i the front element moves to the end.
    void rotate(int i = 1)
    {
        front_ += i;
        if (front_ >= static_cast<int>(_Size)) front_ -= static_cast<int>(_Size);
        else if (front_ < 0) front_ += static_cast<int>(_Size);
    }
};

}  // namespace DROPS

#endif
