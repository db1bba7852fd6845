// drops_b200/hostxx/misc/utils.h -- the slice of DROPS' misc/utils.h that the sparse-solver
// headers need (reference: src/misc/utils.h:131-220 debug classes/Assert, :226-271 valarray
// macros, :369-383 DROPSErrCL, Addr()).  Host side of the B200 drop-in; no CUDA in here.
#ifndef DROPS_UTILS_H
#define DROPS_UTILS_H

#include <cstddef>
#include <iostream>
#include <string>
#include <valarray>
#include <vector>

namespace DROPS
{

typedef unsigned int  Uint;
typedef unsigned long Ulint;
typedef std::vector<size_t> PermutationT;

/// Error class of DROPS: thrown by value, caught in main, reports through what()/handle().
class DROPSErrCL
{
  protected:
    std::string _ErrMesg;

  public:
    DROPSErrCL() {}
    DROPSErrCL(const std::string& mesg) : _ErrMesg(mesg) {}
    virtual ~DROPSErrCL() {}

    virtual std::ostream& what(std::ostream& out) const { return out << _ErrMesg << std::endl; }
    virtual void handle() const { what(std::cerr); std::cerr.flush(); std::abort(); }
    const std::string& message() const { return _ErrMesg; }
};

// Debug classes (bit mask, same names as the reference; only DebugNumericC matters here).
#define DebugContainerC   1
#define DebugRefineEasyC  2
#define DebugRefineHardC  4
#define DebugNumericC     8
#define DebugUnknownsC   16
#define DebugNoReuseC    32
#define DebugParallelC   64
#ifndef DROPSDebugC
#  define DROPSDebugC 0
#endif

template <class E, class A>
inline void _Assert(A assertion, E exc, Uint level = ~0u)
{
    if ((level & DROPSDebugC) && !assertion) throw exc;
}
template <class A>
inline void _Assert(A assertion, const char* msg, Uint level = ~0u)
{
    if ((level & DROPSDebugC) && !assertion) throw DROPSErrCL(msg);
}
#if DROPSDebugC
#  define Assert(a, b, c) ::DROPS::_Assert((a), (b), (c))
#  define Comment(a, b)   do { if ((b) & DROPSDebugC) std::cout << a; } while (0)
#else
#  define Assert(a, b, c) ((void)0)
#  define Comment(a, b)   ((void)0)
#endif

/// Address of the first element of a (possibly const) valarray.
template <typename T>
inline const T* Addr(const std::valarray<T>& x) { return &(const_cast<std::valarray<T>&>(x)[0]); }
template <typename T>
inline T* Addr(std::valarray<T>& x) { return &(x[0]); }

}  // namespace DROPS

#endif
