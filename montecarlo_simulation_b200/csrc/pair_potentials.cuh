// pair_potentials.cuh — device pair energy u(r^2) and virial r.f(r^2) for the six potentials of
// serial_src/potential.cpp:11-186, in the enum order of potential.h:11-18.
//
// The expressions keep the reference's operand order and operand TYPES: the four shape
// parameters are float (potential.h:151,164) and are promoted where the reference promotes them;
// WCA tests the literal r2 < 1.2599 (potential.cpp:62,80); Yukawa hard-codes epsilon = kappa = 1
// and ignores its kappa argument (potential.cpp:95-113); ThermalStar's force multiplies two floats
// before promoting (potential.cpp:183).  These quirks are part of the results contract
// (defect 4).
#pragma once

enum : int { POT_LJ = 0, POT_WCA = 1, POT_YUKAWA = 2, POT_ATHERMAL = 3, POT_THERMAL = 4, POT_IDEAL = 5 };

struct PairParams {
    float f_prime, f_d_prime, kappa, alpha;
};

// energy only
template <int POT>
__device__ __forceinline__ double pair_u(double r2, const PairParams &p) {
    if (POT == POT_LJ) {  // potential.cpp:31-37
        double r2inv = 1.0 / r2;
        double r6inv = r2inv * r2inv * r2inv;
        return 4.0 * r6inv * (r6inv - 1.0);
    } else if (POT == POT_WCA) {  // potential.cpp:59-69
        if (r2 < 1.2599) {
            double r2inv = 1.0 / r2;
            double r6inv = r2inv * r2inv * r2inv;
            return 4.0 * r6inv * (r6inv - 1.0) + 1.0;
        }
        return 0.0;
    } else if (POT == POT_YUKAWA) {  // potential.cpp:95-100
        double r = sqrt(r2);
        return exp(-r) / r;
    } else if (POT == POT_ATHERMAL) {  // potential.cpp:122-133
        double e;
        if (r2 < 1) {
            double r = sqrt(r2);
            e = -log(r) + (double)p.alpha;
        } else {
            e = (double)p.alpha * exp((1 - r2) / 2 / (double)p.alpha);
        }
        return (double)p.f_prime * e;
    } else if (POT == POT_THERMAL) {  // potential.cpp:157-171
        double e1;
        double r = sqrt(r2);
        if (r2 < 1) {
            e1 = -log(r) + (double)p.alpha;
        } else {
            e1 = (double)p.alpha * exp((1 - r2) / 2 / (double)p.alpha);
        }
        double e2 = (double)p.f_d_prime * exp((double)(-p.kappa) * r);
        return (double)p.f_prime * e1 - e2;
    } else {  // Ideal, potential.cpp:11-13
        return 0.0;
    }
}

// energy and virial together (one division / sqrt / exp shared where the formulas allow)
template <int POT>
__device__ __forceinline__ void pair_uw(double r2, const PairParams &p, double &u, double &w) {
    if (POT == POT_LJ) {  // potential.cpp:31-51
        double r2inv = 1.0 / r2;
        double r6inv = r2inv * r2inv * r2inv;
        u = 4.0 * r6inv * (r6inv - 1.0);
        w = 48.0 * r6inv * (r6inv - 0.5);
    } else if (POT == POT_WCA) {  // potential.cpp:59-87
        if (r2 < 1.2599) {
            double r2inv = 1.0 / r2;
            double r6inv = r2inv * r2inv * r2inv;
            u = 4.0 * r6inv * (r6inv - 1.0) + 1.0;
            w = 48.0 * r6inv * (r6inv - 0.5);
        } else {
            u = 0.0;
            w = 0.0;
        }
    } else if (POT == POT_YUKAWA) {  // potential.cpp:95-113
        double r = sqrt(r2);
        double ex = exp(-r);
        u = ex / r;
        w = (1.0 + 1.0 / r) * ex;
    } else if (POT == POT_ATHERMAL) {  // potential.cpp:122-148
        if (r2 < 1) {
            double r = sqrt(r2);
            u = (double)p.f_prime * (-log(r) + (double)p.alpha);
            w = (double)p.f_prime;
        } else {
            double ex = exp((1 - r2) / 2 / (double)p.alpha);
            u = (double)p.f_prime * ((double)p.alpha * ex);
            w = (double)p.f_prime * r2 * ex;
        }
    } else if (POT == POT_THERMAL) {  // potential.cpp:157-188
        double r = sqrt(r2);
        double att = exp((double)(-p.kappa) * r);
        float fk = p.f_d_prime * p.kappa;  // float product, as in the reference
        double f = (double)fk * att * r;
        if (r2 < 1) {
            u = (double)p.f_prime * (-log(r) + (double)p.alpha) - (double)p.f_d_prime * att;
            w = (double)p.f_prime - f;
        } else {
            double ex = exp((1 - r2) / 2 / (double)p.alpha);
            u = (double)p.f_prime * ((double)p.alpha * ex) - (double)p.f_d_prime * att;
            w = (double)p.f_prime * r2 * ex - f;
        }
    } else {
        u = 0.0;
        w = 0.0;
    }
}
