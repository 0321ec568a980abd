// TEST INFRASTRUCTURE ONLY (see oracle/README.md).
// The unmodified reference translation unit src/lib/breakdancer/BreakDancer.cpp, compiled with
// the call `cdf(complement(poisson, readcount))` inside ComputeProbScore (BreakDancer.cpp:65)
// routed through a tracing wrapper, so that a run of the otherwise unmodified reference prints
// every (lambda, k, tail probability) it evaluates to stderr as "POISSON\t<lambda>\t<k>\t<log p>".
// The golden chr21 rows all print score 99 (the cap), so this is how their raw log-probabilities
// are pinned (tests/golden/chr21_poisson_trace.tsv). Built by oracle/build_ref.sh and linked with
// the reference's other objects into oracle/_ref/breakdancer-max-trace. No reference source is
// copied: the TU is included from where it lies.
#include <boost/array.hpp>
#include <boost/bind.hpp>
#include <boost/lexical_cast.hpp>
#include <boost/math/distributions/chi_squared.hpp>
#include <boost/math/distributions/poisson.hpp>
#include <boost/math/distributions/complement.hpp>
#include <boost/ref.hpp>
#include <boost/format.hpp>
#include <boost/function.hpp>
#include <boost/chrono.hpp>
#include <boost/unordered_map.hpp>
#include <boost/unordered_set.hpp>
#include <boost/scoped_ptr.hpp>
#include <boost/shared_ptr.hpp>
#include <boost/noncopyable.hpp>
#include <boost/iterator/filter_iterator.hpp>
#include <boost/range/iterator_range.hpp>
#include <boost/range/algorithm/remove_copy_if.hpp>
#include <boost/range/algorithm/for_each.hpp>
#include <boost/serialization/array.hpp>
#include <boost/serialization/string.hpp>
#include <boost/serialization/vector.hpp>
#include <boost/serialization/nvp.hpp>
#include <boost/container/flat_map.hpp>
#include <cstdio>
#include <cmath>

namespace bd_trace {
template <class Dist, class K>
double traced(boost::math::complemented2_type<Dist, K> const& c) {
    double p = boost::math::cdf(c);
    std::fprintf(stderr, "POISSON\t%.17g\t%d\t%.17g\n", (double)c.dist.mean(), (int)c.param, std::log(p));
    return p;
}
template <class K>
double traced(boost::math::complemented2_type<boost::math::chi_squared, K> const& c) {
    return boost::math::cdf(c);
}
}  // namespace bd_trace

#define cdf(X) bd_trace::traced(X)
#include "breakdancer/BreakDancer.cpp"
