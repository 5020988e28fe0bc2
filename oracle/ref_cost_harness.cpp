// TEST INFRASTRUCTURE ONLY (oracle/): C entry points around the reference's own cost-model pieces — the global
// EDT_Matrix (planner.cpp:37), ValidityChecker::isValid / clearance (:587-632), ClearanceObjective::stateCost (:655-669),
// shortrisky / longsafe (:677-690) — which precede this file in the translation unit assembled by oracle/Makefile
// (target ref_cost -> oracle/_ref/libref_cost.so).  Used by tests/test_cost_reference_pieces.py to pin
// oracle/cost_oracle.c's per-sample (cell, valid, 1/clearance) and the objective weights to reference-EXECUTED code.
namespace {
struct Rig {
    ob::SpaceInformationPtr si;
    ob::OptimizationObjectivePtr clear;
    Rig() : si(new ob::SpaceInformation()) {
        si->setStateValidityChecker(ob::StateValidityCheckerPtr(new ValidityChecker(si)));   // planner.cpp:699
        clear = getClearanceObjective(si);                                                     // planner.cpp:671-674
    }
};
Rig& rig() { static Rig r; return r; }
struct Point : ob::RealVectorStateSpace::StateType {
    double xy[2];
    Point(double x, double y) { xy[0] = x; xy[1] = y; values = xy; }
};
}  // namespace

extern "C" {
// E given row-major [row = y][col = x]; stored into the reference's Eigen matrix as EDT_Matrix(row, col)
void ref_cost_set_map(const double* e, int rows, int cols) {
    EDT_Matrix.resize(rows, cols);
    for (int r = 0; r < rows; ++r)
        for (int c = 0; c < cols; ++c) EDT_Matrix(r, c) = e[(size_t)r * cols + c];
}
int ref_cost_is_valid(double x, double y) { Point p(x, y); return rig().si->getStateValidityChecker()->isValid(&p) ? 1 : 0; }
double ref_cost_clearance(double x, double y) { Point p(x, y); return rig().si->getStateValidityChecker()->clearance(&p); }
double ref_cost_state_cost(double x, double y) { Point p(x, y); return rig().clear->stateCost(&p).value(); }
// kind 0 = shortrisky, 1 = longsafe: out2 = (weight of the path-length objective, weight of the clearance objective);
// returns 1 when the clearance term is a StateCostIntegralObjective with motion-cost interpolation enabled (:651)
int ref_cost_weights(int kind, double* out2) {
    ob::OptimizationObjectivePtr o = kind == 0 ? shortrisky(rig().si) : longsafe(rig().si);
    const ob::MultiOptimizationObjective* m = dynamic_cast<const ob::MultiOptimizationObjective*>(o.get());
    if (!m || m->getObjectiveCount() != 2) return -1;
    int interp = -1;
    out2[0] = out2[1] = 0.0;
    for (std::size_t i = 0; i < 2; ++i) {
        const ob::StateCostIntegralObjective* sc = dynamic_cast<const ob::StateCostIntegralObjective*>(m->getObjective(i).get());
        if (sc) { out2[1] = m->getObjectiveWeight(i); interp = sc->isMotionCostInterpolationEnabled() ? 1 : 0; }
        else if (dynamic_cast<const ob::PathLengthOptimizationObjective*>(m->getObjective(i).get())) out2[0] = m->getObjectiveWeight(i);
    }
    return interp;
}
double ref_cost_threshold_path_length() { return getThresholdPathLengthObj(rig().si)->getCostThreshold().value(); }   // :641-647
}
