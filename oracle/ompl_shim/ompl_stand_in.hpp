// TEST INFRASTRUCTURE ONLY (oracle/): a minimal stand-in for the OMPL types the reference's cost-model pieces are written
// against (sample_based_optimisation_based_path_planner.cpp:587-690: ValidityChecker, ClearanceObjective, shortrisky,
// longsafe).  OMPL is not installed in this image and not vendored by the reference (find_package(ompl REQUIRED), no
// version pinned).  Only what those ~100 lines touch exists here: State / RealVectorStateSpace::StateType with `values`,
// StateValidityChecker (isValid / clearance), SpaceInformation holding the checker, Cost, OptimizationObjective with a
// cost threshold, PathLengthOptimizationObjective, StateCostIntegralObjective (constructor flag + stateCost hook), and
// the `w * objective + ...` operators that build a weighted MultiOptimizationObjective.  No planner, no interpolation,
// no motion cost: the path-integral quadrature is NOT reference code and stays DECLARED in oracle/cost_oracle.c.
#pragma once
#include <cmath>
#include <iostream>
#include <memory>
#include <utility>
#include <vector>

namespace ompl {
namespace base {

struct State {
    virtual ~State() {}
    template <class T> const T* as() const { return static_cast<const T*>(this); }
    template <class T> T* as() { return static_cast<T*>(this); }
};

class RealVectorStateSpace {
public:
    struct StateType : public State {
        double* values;
        StateType() : values(0) {}
    };
};

class SpaceInformation;
typedef std::shared_ptr<SpaceInformation> SpaceInformationPtr;

class StateValidityChecker {
public:
    explicit StateValidityChecker(const SpaceInformationPtr& si) : si_(si.get()) {}
    virtual ~StateValidityChecker() {}
    virtual bool isValid(const State* state) const = 0;
    virtual double clearance(const State*) const { return 0.0; }
protected:
    SpaceInformation* si_;
};
typedef std::shared_ptr<StateValidityChecker> StateValidityCheckerPtr;

class SpaceInformation {
public:
    void setStateValidityChecker(const StateValidityCheckerPtr& c) { checker_ = c; }
    const StateValidityCheckerPtr& getStateValidityChecker() const { return checker_; }
private:
    StateValidityCheckerPtr checker_;
};

class Cost {
public:
    explicit Cost(double v = 0.0) : v_(v) {}
    double value() const { return v_; }
private:
    double v_;
};
inline std::ostream& operator<<(std::ostream& os, const Cost& c) { return os << c.value(); }

class OptimizationObjective {
public:
    explicit OptimizationObjective(const SpaceInformationPtr& si) : si_(si), threshold_(0.0) {}
    virtual ~OptimizationObjective() {}
    void setCostThreshold(Cost c) { threshold_ = c; }
    Cost getCostThreshold() const { return threshold_; }
    virtual Cost stateCost(const State*) const { return Cost(1.0); }
    const SpaceInformationPtr& getSpaceInformation() const { return si_; }
protected:
    SpaceInformationPtr si_;
    Cost threshold_;
};
typedef std::shared_ptr<OptimizationObjective> OptimizationObjectivePtr;

class PathLengthOptimizationObjective : public OptimizationObjective {
public:
    explicit PathLengthOptimizationObjective(const SpaceInformationPtr& si) : OptimizationObjective(si) {}
};

class StateCostIntegralObjective : public OptimizationObjective {
public:
    StateCostIntegralObjective(const SpaceInformationPtr& si, bool enableMotionCostInterpolation = false)
        : OptimizationObjective(si), interpolateMotionCost_(enableMotionCostInterpolation) {}
    bool isMotionCostInterpolationEnabled() const { return interpolateMotionCost_; }
protected:
    bool interpolateMotionCost_;
};

// weighted sum of objectives, as built by `w1 * a + w2 * b`
class MultiOptimizationObjective : public OptimizationObjective {
public:
    explicit MultiOptimizationObjective(const SpaceInformationPtr& si) : OptimizationObjective(si) {}
    void addObjective(const OptimizationObjectivePtr& o, double w) { parts_.push_back(std::make_pair(o, w)); }
    std::size_t getObjectiveCount() const { return parts_.size(); }
    const OptimizationObjectivePtr& getObjective(std::size_t i) const { return parts_[i].first; }
    double getObjectiveWeight(std::size_t i) const { return parts_[i].second; }
private:
    std::vector<std::pair<OptimizationObjectivePtr, double> > parts_;
};

inline OptimizationObjectivePtr operator*(double w, const OptimizationObjectivePtr& a) {
    MultiOptimizationObjective* m = new MultiOptimizationObjective(a->getSpaceInformation());
    m->addObjective(a, w);
    return OptimizationObjectivePtr(m);
}
inline OptimizationObjectivePtr operator+(const OptimizationObjectivePtr& a, const OptimizationObjectivePtr& b) {
    MultiOptimizationObjective* m = new MultiOptimizationObjective(a->getSpaceInformation());
    const OptimizationObjectivePtr two[2] = {a, b};
    for (int k = 0; k < 2; ++k) {
        const MultiOptimizationObjective* mk = dynamic_cast<const MultiOptimizationObjective*>(two[k].get());
        if (mk) for (std::size_t i = 0; i < mk->getObjectiveCount(); ++i) m->addObjective(mk->getObjective(i), mk->getObjectiveWeight(i));
        else m->addObjective(two[k], 1.0);
    }
    return OptimizationObjectivePtr(m);
}

}  // namespace base
namespace geometric {}
}  // namespace ompl
