// lmcma.hpp - drop-in for the reference header of the same name (lmcma_path_planner/src/lmcma.hpp): put this directory on
// the include path INSTEAD of the reference's src/ and link -llmcma_b200; a translation unit written against the
// reference header (its own example_lmcma.cpp, for one: tests/test_capi_cpu.py compiles it unchanged) then builds
// against the B200 library.  Everything is declared in include/lmcma_b200.hpp; this file only lifts the reference's
// global names out of the namespace.  SepCMA / CMAChol are not provided (out of scope: never instantiated by the
// reference).
#ifndef LMCMA_B200_DROP_IN_LMCMA_HPP
#define LMCMA_B200_DROP_IN_LMCMA_HPP
#include "lmcma_b200.hpp"

using lmcma_b200::sortedvals;
using lmcma_b200::random_t;
using lmcma_b200::random_exit;
using lmcma_b200::random_Start;
using lmcma_b200::random_init;
using lmcma_b200::random_Uniform;
using lmcma_b200::random_Gauss;
using lmcma_b200::compare;
using lmcma_b200::myqsort;
using lmcma_b200::CMABase;
using lmcma_b200::LMCMA;
using lmcma_b200::covariance;
using lmcma_b200::differentiationMatrix;
using lmcma_b200::invert;
using lmcma_b200::cholesky;
using lmcma_b200::applyCovL;
#endif
