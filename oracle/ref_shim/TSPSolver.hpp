// oracle/ref_shim/TSPSolver.hpp -- TEST INFRASTRUCTURE.  Shadows the reference's TSP solvers (nlopt-based, not on the path) so
// that tests/CameraPathGen.cpp compiles; only reached from Planner::run_* methods that nothing here calls.
#pragma once
#include <cstdlib>
#include <vector>
namespace TSP {
struct SolverStub {
    std::vector<int> path_;
    explicit SolverStub(std::vector<std::vector<int>>&) {}
    void solve() { std::abort(); }
    int getPathLength() { return 0; }
};
typedef SolverStub NearestNeighborSearch;
typedef SolverStub TwoOptSearch;
}
