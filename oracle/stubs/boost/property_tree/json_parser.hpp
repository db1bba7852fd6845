// Empty stand-in: misc/utils.h:849-851 of the reference includes this Boost header,
// but nothing on the sparse-solver hot path uses it. Test infrastructure only.
