/**
 * ani_estimator.hpp -- drop-in for the reference's src/ani_estimator.hpp:13-14.
 * Host double arithmetic on integer counts (src/ani_estimation.cpp:24-42); implemented by
 * sks_containment / sks_binomial_estimator of the C ABI (include/sks.h).
 */
#ifndef SKS_ANI_ESTIMATOR_HPP
#define SKS_ANI_ESTIMATOR_HPP
#include <cmath>
#include "logging.hpp"

/** |A n B| / |A|, 0 when the intersection is empty (src/ani_estimation.cpp:24-28). */
double containment(int intersection, int set_size);
/** containment^(1 / kmer_num_ones), 0 for containment <= 0 (src/ani_estimation.cpp:38-42). */
double binomial_estimator(double containment, int kmer_num_ones);
#endif
