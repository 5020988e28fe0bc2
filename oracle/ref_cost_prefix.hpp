// TEST INFRASTRUCTURE ONLY (oracle/): what precedes the reference's cost-model lines in the translation unit that
// oracle/Makefile assembles on the fly (prefix + `sed -n` of the reference file where it lies + oracle/ref_cost_harness.cpp,
// piped to g++; the extracted lines are never written into this repository).
#include <cmath>
#include <cstdlib>
#include <iostream>
#include <eigen3/Eigen/Dense>             // oracle/eigen_shim: MatrixXd with (row, col) access, column-major like Eigen
#include "ompl_stand_in.hpp"              // oracle/ompl_shim
using namespace std;
