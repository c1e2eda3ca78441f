// TimeMeasure.h -- per-stage accumulators the loop fills (reference: icp-variants/TimeMeasure.h:7-62).
// The device loop fills them from CUDA events (seconds): matchingTime covers transform + search +
// weighting + rejection (one fused kernel), solverTime the residual/Jacobian reduction + solve + pose
// update, weighingTime / rejectionTime stay 0 because those stages are fused into matching.
#pragma once
#include <iostream>

class TimeMeasure {
public:
    double selectionTime, matchingTime, weighingTime, rejectionTime, solverTime, convergenceTime, indexTime;
    unsigned int* nIterations;
    TimeMeasure() : selectionTime(0), matchingTime(0), weighingTime(0), rejectionTime(0), solverTime(0), convergenceTime(0), indexTime(0), nIterations(nullptr) {}
    void calculateIterationTime() {
        const double n = (nIterations && *nIterations) ? (double)*nIterations : 1.0;
        std::cout << "Convergence time = " << convergenceTime << " s\nTime taken for each step (average):\n"
                  << "\t [*] Selection time = " << selectionTime / n << " s \n"
                  << "\t [*] Matching time = " << matchingTime / n << " s per iteration\n"
                  << "\t [*] Weighing time = " << weighingTime / n << " s per iteration\n"
                  << "\t [*] Rejection time = " << rejectionTime / n << " s per iteration\n"
                  << "\t [*] Minimization (one icp step) time = " << solverTime / n << " s per iteration\n";
    }
};
