// TEST INFRASTRUCTURE ONLY (see oracle/README.md).
// Known-answer generator for the Poisson tail used by the reference's ComputeProbScore
// (/root/reference/src/lib/breakdancer/BreakDancer.cpp:62-68): evaluates, with the reference's
// own vendored boost 1.54 headers, log(cdf(complement(poisson_distribution<double>(lambda), k)))
// for "lambda k" pairs read from stdin and prints "lambda k logp" with 17 significant digits.
// Built by oracle/build_ref.sh into oracle/_ref/score_ref; its output is committed as
// tests/golden/poisson_logp.tsv by tests/golden/make_golden.py.
#include <boost/math/distributions/poisson.hpp>
#include <boost/math/distributions/chi_squared.hpp>
#include <cstdio>
#include <cmath>
int main(int argc, char** argv) {
    using namespace boost::math;
    bool fisher = argc > 1 && argv[1][0] == 'f';
    double a, b;
    if (!fisher) {
        // rows: lambda k
        while (std::scanf("%lf %lf", &a, &b) == 2) {
            poisson_distribution<double> pois(a);
            double lp = std::log(cdf(complement(pois, (int)b)));
            std::printf("%.17g\t%d\t%.17g\n", a, (int)b, lp);
        }
    } else {
        // rows: ndf x  -> log(cdf(complement(chi_squared(ndf), x)))   (BreakDancer.cpp:73-76)
        while (std::scanf("%lf %lf", &a, &b) == 2) {
            chi_squared chisq(a);
            double p = cdf(complement(chisq, b));
            std::printf("%.17g\t%.17g\t%.17g\n", a, b, p);
        }
    }
    return 0;
}
