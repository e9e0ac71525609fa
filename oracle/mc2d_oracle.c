/*
 * mc2d_oracle.c — CPU ORACLE (test infrastructure, NOT product code).
 *
 * Plain-C restatement of the reference's serial 2D Monte Carlo hot path; see
 * mc2d_oracle.h for the contract and the parity status (pinned against the
 * reference's known-answer tests, against the reference compiled in place
 * (oracle/_ref), and against the sanity_check golden CSV).
 *
 * The arithmetic below follows the reference expression by expression (operand
 * order, float-vs-double operand types, division-vs-reciprocal) so that results
 * are bit-identical to the reference built with g++ -O3 on x86-64 (no FMA
 * contraction: build this file with -ffp-contract=off).
 */
#include "mc2d_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------- */
/* pair functions — serial_src/potential.cpp                                   */
/* ------------------------------------------------------------------------- */

/* potential.cpp:31-37 */
static double lj_u(double r2) {
    const double epsilon = 1.0, sigma = 1.0;
    double r2inv = (sigma * sigma) / r2;
    double r6inv = r2inv * r2inv * r2inv;
    return 4.0 * epsilon * r6inv * (r6inv - 1.0);
}
/* potential.cpp:45-51 */
static double lj_w(double r2) {
    const double epsilon = 1.0, sigma = 1.0;
    double r2inv = (sigma * sigma) / r2;
    double r6inv = r2inv * r2inv * r2inv;
    return 48.0 * epsilon * r6inv * (r6inv - 0.5);
}
/* potential.cpp:59-69 — the cutoff is the literal 1.2599, not 2^(1/3) */
static double wca_u(double r2) {
    const double epsilon = 1.0, sigma = 1.0;
    if (r2 < 1.2599) {
        double r2inv = sigma * sigma / r2;
        double r6inv = r2inv * r2inv * r2inv;
        return 4.0 * epsilon * r6inv * (r6inv - 1.0) + epsilon;
    }
    return 0.0;
}
/* potential.cpp:77-87 */
static double wca_w(double r2) {
    const double epsilon = 1.0, sigma = 1.0;
    if (r2 < 1.2599) {
        double r2inv = sigma * sigma / r2;
        double r6inv = r2inv * r2inv * r2inv;
        return 48.0 * epsilon * r6inv * (r6inv - 0.5);
    }
    return 0.0;
}
/* potential.cpp:95-100 — epsilon = kappa = 1 hard-coded; the kappa argument is ignored */
static double yukawa_u(double r2) {
    const double epsilon = 1.0, kappa = 1.0;
    double r = sqrt(r2);
    return epsilon * exp(-kappa * r) / r;
}
/* potential.cpp:108-113 */
static double yukawa_w(double r2) {
    const double epsilon = 1.0, kappa = 1.0;
    double r = sqrt(r2);
    return epsilon * (kappa + 1.0 / r) * exp(-kappa * r);
}
/* potential.cpp:122-133 */
static double athermal_u(double r2, float f_dependant, float alpha) {
    double e;
    if (r2 < 1) {
        double r = sqrt(r2);
        e = -log(r) + alpha;
    } else {
        e = alpha * exp((1 - r2) / 2 / alpha);
    }
    return f_dependant * e;
}
/* potential.cpp:142-148 */
static double athermal_w(double r2, float f_dependant, float alpha) {
    if (r2 < 1) {
        return f_dependant;
    }
    return f_dependant * r2 * exp((1 - r2) / 2 / alpha);
}
/* potential.cpp:157-171 */
static double thermal_u(double r2, float f_dependant, float f_dependent_2, float kappa, float alpha) {
    double e1;
    double r = sqrt(r2);
    if (r2 < 1) {
        e1 = -log(r) + alpha;
    } else {
        e1 = alpha * exp((1 - r2) / 2 / alpha);
    }
    double e2;
    e2 = f_dependent_2 * exp(-kappa * r);
    return f_dependant * e1 - e2;
}
/* potential.cpp:180-188 — note f_dependent_2 * kappa is a float*float product */
static double thermal_w(double r2, float f_dependant, float f_dependent_2, float kappa, float alpha) {
    double f;
    double r = sqrt(r2);
    f = f_dependent_2 * kappa * exp(-kappa * r) * r;
    if (r2 < 1) {
        return f_dependant - f;
    }
    return f_dependant * r2 * exp((1 - r2) / 2 / alpha) - f;
}

/* potential.cpp:212-229 */
double orc_pair_potential(double r2, int type, float f_prime, float f_d_prime, float kappa, float alpha) {
    switch (type) {
    case ORC_LJ: return lj_u(r2);
    case ORC_WCA: return wca_u(r2);
    case ORC_YUKAWA: return yukawa_u(r2);
    case ORC_ATHERMAL_STAR: return athermal_u(r2, f_prime, alpha);
    case ORC_THERMAL_STAR: return thermal_u(r2, f_prime, f_d_prime, kappa, alpha);
    case ORC_IDEAL: return 0.0; /* potential.cpp:11-13 */
    default: return NAN;        /* reference throws std::invalid_argument */
    }
}
/* potential.cpp:239-256 */
double orc_pair_force(double r2, int type, float f_prime, float f_d_prime, float kappa, float alpha) {
    switch (type) {
    case ORC_LJ: return lj_w(r2);
    case ORC_WCA: return wca_w(r2);
    case ORC_YUKAWA: return yukawa_w(r2);
    case ORC_ATHERMAL_STAR: return athermal_w(r2, f_prime, alpha);
    case ORC_THERMAL_STAR: return thermal_w(r2, f_prime, f_d_prime, kappa, alpha);
    case ORC_IDEAL: return 0.0; /* potential.cpp:20-22 */
    default: return NAN;
    }
}
/* potential.cpp:194-202 */
int orc_select_potential(const char *name) {
    if (!strcmp(name, "LennardJones")) return ORC_LJ;
    if (!strcmp(name, "WCA")) return ORC_WCA;
    if (!strcmp(name, "Yukawa")) return ORC_YUKAWA;
    if (!strcmp(name, "AthermalStar")) return ORC_ATHERMAL_STAR;
    if (!strcmp(name, "ThermalStar")) return ORC_THERMAL_STAR;
    if (!strcmp(name, "Ideal")) return ORC_IDEAL;
    return -1;
}

/* thermodynamic_calculator.cpp:19-21 */
void orc_calc_params(double f, double A0, double kappa, double *f_prime, double *f_d_prime) {
    *f_prime = (2.0 + 9.0 * f * f) / 24.0;
    *f_d_prime = A0 * f * f / kappa;
}
/* thermodynamic_calculator.cpp:5-26 */
void orc_calc_init(orc_calc *c, double T, double press, int type, double rcut, double mu, double f,
                   double alpha, double A0, double kappa) {
    c->T = T;
    c->press = press;
    c->rcut = rcut;
    c->r2cut = rcut * rcut;
    c->type = type;
    orc_calc_params(f, A0, kappa, &c->f_prime, &c->f_d_prime);
    c->alpha = alpha;
    c->kappa = kappa;
    c->z = mu;
}
static double calc_u(const orc_calc *c, double r2) {
    /* doubles narrowed to float at the call, as potential.h:151 forces */
    return orc_pair_potential(r2, c->type, (float)c->f_prime, (float)c->f_d_prime, (float)c->kappa, (float)c->alpha);
}
static double calc_w(const orc_calc *c, double r2) {
    return orc_pair_force(r2, c->type, (float)c->f_prime, (float)c->f_d_prime, (float)c->kappa, (float)c->alpha);
}

/* ------------------------------------------------------------------------- */
/* box — serial_src/initial.cpp                                                */
/* ------------------------------------------------------------------------- */

/* initial.cpp:18-19 (invL precomputed), :22-25 */
void orc_apply_pbc(double Lx, double Ly, double *x, double *y) {
    double invLx = 1.0 / Lx, invLy = 1.0 / Ly;
    *x -= Lx * floor(*x * invLx);
    *y -= Ly * floor(*y * invLy);
}
/* initial.cpp:43-51 */
double orc_min_image_r2(double Lx, double Ly, double x1, double y1, double x2, double y2) {
    double invLx = 1.0 / Lx, invLy = 1.0 / Ly;
    double dx = x1 - x2;
    double dy = y1 - y2;
    dx -= Lx * round(dx * invLx);
    dy -= Ly * round(dy * invLy);
    return (dx * dx + dy * dy);
}

/* ------------------------------------------------------------------------- */
/* cell list — serial_src/cell_list.cpp                                        */
/* ------------------------------------------------------------------------- */

struct orc_celllist {
    double Lx, Ly, rcut, rcutsq;
    int nx, ny;
    double cell_dx, cell_dy;
    /* buckets in CSR form; within a bucket particles keep push_back (index) order */
    int *start;         /* nx*ny + 1 */
    int *items;         /* N */
    int *particle_cell; /* N */
    int N, cap;
};

/* cell_list.cpp:5-15 */
orc_celllist *orc_cl_create(double Lx, double Ly, double rcut) {
    orc_celllist *cl = (orc_celllist *)calloc(1, sizeof(*cl));
    cl->Lx = Lx;
    cl->Ly = Ly;
    cl->rcut = rcut;
    cl->rcutsq = rcut * rcut;
    cl->nx = (int)(Lx / rcut);
    cl->ny = (int)(Ly / rcut);
    if (cl->nx < 1) cl->nx = 1;
    if (cl->ny < 1) cl->ny = 1;
    cl->cell_dx = Lx / cl->nx;
    cl->cell_dy = Ly / cl->ny;
    cl->start = (int *)calloc((size_t)cl->nx * cl->ny + 1, sizeof(int));
    return cl;
}
void orc_cl_destroy(orc_celllist *cl) {
    if (!cl) return;
    free(cl->start);
    free(cl->items);
    free(cl->particle_cell);
    free(cl);
}
void orc_cl_dims(const orc_celllist *cl, int *nx, int *ny, double *cell_dx, double *cell_dy) {
    if (nx) *nx = cl->nx;
    if (ny) *ny = cl->ny;
    if (cell_dx) *cell_dx = cl->cell_dx;
    if (cell_dy) *cell_dy = cl->cell_dy;
}
/* cell_list.cpp:103-109 */
int orc_cl_cell_index(const orc_celllist *cl, double x, double y) {
    int cx = (int)floor(x / cl->cell_dx);
    int cy = (int)floor(y / cl->cell_dy);
    cx = (cx % cl->nx + cl->nx) % cl->nx;
    cy = (cy % cl->ny + cl->ny) % cl->ny;
    return cx + cy * cl->nx;
}
/* cell_list.cpp:28-41 — counting sort that keeps index order inside each bucket */
void orc_cl_build(orc_celllist *cl, const double *xy, int N) {
    int nc = cl->nx * cl->ny;
    if (N > cl->cap) {
        cl->cap = N + N / 2 + 16;
        cl->items = (int *)realloc(cl->items, sizeof(int) * (size_t)cl->cap);
        cl->particle_cell = (int *)realloc(cl->particle_cell, sizeof(int) * (size_t)cl->cap);
    }
    cl->N = N;
    memset(cl->start, 0, sizeof(int) * ((size_t)nc + 1));
    for (int i = 0; i < N; ++i) {
        int idx = orc_cl_cell_index(cl, xy[2 * i], xy[2 * i + 1]);
        cl->particle_cell[i] = idx;
        cl->start[idx + 1]++;
    }
    for (int c = 0; c < nc; ++c) cl->start[c + 1] += cl->start[c];
    int *fill = (int *)malloc(sizeof(int) * (size_t)(nc > 0 ? nc : 1));
    memcpy(fill, cl->start, sizeof(int) * (size_t)nc);
    for (int i = 0; i < N; ++i) cl->items[fill[cl->particle_cell[i]]++] = i;
    free(fill);
}
/* cell_list.cpp:43-70 */
int orc_cl_neighbors(const orc_celllist *cl, int idx, const double *xy, int *out_j, double *out_r2, int cap) {
    int n = 0;
    int ci = cl->particle_cell[idx];
    int cx = ci % cl->nx;
    int cy = ci / cl->nx;
    for (int ddx = -1; ddx <= 1; ++ddx) {
        for (int ddy = -1; ddy <= 1; ++ddy) {
            int ncx = (cx + ddx + cl->nx) % cl->nx;
            int ncy = (cy + ddy + cl->ny) % cl->ny;
            int nci = ncx + ncy * cl->nx;
            for (int k = cl->start[nci]; k < cl->start[nci + 1]; ++k) {
                int j = cl->items[k];
                if (j == idx) continue;
                double dx = xy[2 * j] - xy[2 * idx];
                double dy = xy[2 * j + 1] - xy[2 * idx + 1];
                dx -= cl->Lx * round(dx / cl->Lx);
                dy -= cl->Ly * round(dy / cl->Ly);
                double r_sq = dx * dx + dy * dy;
                if (r_sq <= cl->rcutsq) {
                    if (n < cap) {
                        if (out_j) out_j[n] = j;
                        if (out_r2) out_r2[n] = r_sq;
                    }
                    ++n;
                }
            }
        }
    }
    return n;
}
/* cell_list.cpp:72-101 */
int orc_cl_neighbors_point(const orc_celllist *cl, double x, double y, const double *xy, int *out_j,
                           double *out_r2, int cap) {
    int n = 0;
    int cx = (int)floor(x / cl->cell_dx);
    int cy = (int)floor(y / cl->cell_dy);
    cx = (cx % cl->nx + cl->nx) % cl->nx;
    cy = (cy % cl->ny + cl->ny) % cl->ny;
    for (int ddx = -1; ddx <= 1; ++ddx) {
        for (int ddy = -1; ddy <= 1; ++ddy) {
            int ncx = (cx + ddx + cl->nx) % cl->nx;
            int ncy = (cy + ddy + cl->ny) % cl->ny;
            int nci = ncx + ncy * cl->nx;
            for (int k = cl->start[nci]; k < cl->start[nci + 1]; ++k) {
                int j = cl->items[k];
                double dx = xy[2 * j] - x;
                double dy = xy[2 * j + 1] - y;
                dx -= cl->Lx * round(dx / cl->Lx);
                dy -= cl->Ly * round(dy / cl->Ly);
                double r_sq = dx * dx + dy * dy;
                if (r_sq <= cl->rcutsq) {
                    if (n < cap) {
                        if (out_j) out_j[n] = j;
                        if (out_r2) out_r2[n] = r_sq;
                    }
                    ++n;
                }
            }
        }
    }
    return n;
}

/* ------------------------------------------------------------------------- */
/* observables — serial_src/thermodynamic_calculator.cpp                       */
/* ------------------------------------------------------------------------- */

/* :56-77 */
double orc_total_energy(const orc_calc *c, double Lx, double Ly, const double *xy, int N) {
    double total = 0.0;
    for (int i = 0; i < N; ++i)
        for (int j = i + 1; j < N; ++j) {
            double r2 = orc_min_image_r2(Lx, Ly, xy[2 * i], xy[2 * i + 1], xy[2 * j], xy[2 * j + 1]);
            if (r2 <= c->r2cut) total += calc_u(c, r2);
        }
    return total;
}
/* :80-101 */
double orc_total_virial(const orc_calc *c, double Lx, double Ly, const double *xy, int N) {
    double virial = 0.0;
    for (int i = 0; i < N; ++i)
        for (int j = i + 1; j < N; ++j) {
            double r2 = orc_min_image_r2(Lx, Ly, xy[2 * i], xy[2 * i + 1], xy[2 * j], xy[2 * j + 1]);
            if (r2 <= c->r2cut) virial += calc_w(c, r2);
        }
    return virial;
}
/* :104-111 */
double orc_pressure(const orc_calc *c, double Lx, double Ly, const double *xy, int N) {
    double V = Lx * Ly;
    double rho = (double)N / V;
    double W = orc_total_virial(c, Lx, Ly, xy, N);
    return rho * c->T + W / (2.0 * V);
}

/* sum over a neighbour scan without materialising the list; same visiting order as
 * computeLocalEnergy over getNeighbors (thermodynamic_calculator.h:103-121) */
static double scan_sum(const orc_celllist *cl, const orc_calc *c, const double *xy, int idx, int virial) {
    double acc = 0.0;
    int ci = cl->particle_cell[idx];
    int cx = ci % cl->nx;
    int cy = ci / cl->nx;
    for (int ddx = -1; ddx <= 1; ++ddx)
        for (int ddy = -1; ddy <= 1; ++ddy) {
            int ncx = (cx + ddx + cl->nx) % cl->nx;
            int ncy = (cy + ddy + cl->ny) % cl->ny;
            int nci = ncx + ncy * cl->nx;
            for (int k = cl->start[nci]; k < cl->start[nci + 1]; ++k) {
                int j = cl->items[k];
                if (j == idx) continue;
                double dx = xy[2 * j] - xy[2 * idx];
                double dy = xy[2 * j + 1] - xy[2 * idx + 1];
                dx -= cl->Lx * round(dx / cl->Lx);
                dy -= cl->Ly * round(dy / cl->Ly);
                double r_sq = dx * dx + dy * dy;
                if (r_sq <= cl->rcutsq) acc += virial ? calc_w(c, r_sq) : calc_u(c, r_sq);
            }
        }
    return acc;
}
static double scan_sum_point(const orc_celllist *cl, const orc_calc *c, const double *xy, double x, double y,
                             int exclude) {
    double acc = 0.0;
    int cx = (int)floor(x / cl->cell_dx);
    int cy = (int)floor(y / cl->cell_dy);
    cx = (cx % cl->nx + cl->nx) % cl->nx;
    cy = (cy % cl->ny + cl->ny) % cl->ny;
    for (int ddx = -1; ddx <= 1; ++ddx)
        for (int ddy = -1; ddy <= 1; ++ddy) {
            int ncx = (cx + ddx + cl->nx) % cl->nx;
            int ncy = (cy + ddy + cl->ny) % cl->ny;
            int nci = ncx + ncy * cl->nx;
            for (int k = cl->start[nci]; k < cl->start[nci + 1]; ++k) {
                int j = cl->items[k];
                if (j == exclude) continue;
                double dx = xy[2 * j] - x;
                double dy = xy[2 * j + 1] - y;
                dx -= cl->Lx * round(dx / cl->Lx);
                dy -= cl->Ly * round(dy / cl->Ly);
                double r_sq = dx * dx + dy * dy;
                if (r_sq <= cl->rcutsq) acc += calc_u(c, r_sq);
            }
        }
    return acc;
}

/* :194-211 — a fresh CellList per call, directed pairs summed then halved */
double orc_total_energy_celllist(const orc_calc *c, double Lx, double Ly, const double *xy, int N) {
    orc_celllist *cl = orc_cl_create(Lx, Ly, c->rcut);
    orc_cl_build(cl, xy, N);
    double U = 0.0;
    for (int i = 0; i < N; ++i) {
        /* the reference adds pair by pair into U (not a per-particle subtotal) */
        int ci = cl->particle_cell[i];
        int cx = ci % cl->nx, cy = ci / cl->nx;
        for (int ddx = -1; ddx <= 1; ++ddx)
            for (int ddy = -1; ddy <= 1; ++ddy) {
                int nci = (cx + ddx + cl->nx) % cl->nx + ((cy + ddy + cl->ny) % cl->ny) * cl->nx;
                for (int k = cl->start[nci]; k < cl->start[nci + 1]; ++k) {
                    int j = cl->items[k];
                    if (j == i) continue;
                    double dx = xy[2 * j] - xy[2 * i];
                    double dy = xy[2 * j + 1] - xy[2 * i + 1];
                    dx -= cl->Lx * round(dx / cl->Lx);
                    dy -= cl->Ly * round(dy / cl->Ly);
                    double r_sq = dx * dx + dy * dy;
                    if (r_sq <= cl->rcutsq) U += calc_u(c, r_sq);
                }
            }
    }
    orc_cl_destroy(cl);
    return 0.5 * U;
}
/* :215-232 */
double orc_total_virial_celllist(const orc_calc *c, double Lx, double Ly, const double *xy, int N) {
    orc_celllist *cl = orc_cl_create(Lx, Ly, c->rcut);
    orc_cl_build(cl, xy, N);
    double W = 0.0;
    for (int i = 0; i < N; ++i) {
        int ci = cl->particle_cell[i];
        int cx = ci % cl->nx, cy = ci / cl->nx;
        for (int ddx = -1; ddx <= 1; ++ddx)
            for (int ddy = -1; ddy <= 1; ++ddy) {
                int nci = (cx + ddx + cl->nx) % cl->nx + ((cy + ddy + cl->ny) % cl->ny) * cl->nx;
                for (int k = cl->start[nci]; k < cl->start[nci + 1]; ++k) {
                    int j = cl->items[k];
                    if (j == i) continue;
                    double dx = xy[2 * j] - xy[2 * i];
                    double dy = xy[2 * j + 1] - xy[2 * i + 1];
                    dx -= cl->Lx * round(dx / cl->Lx);
                    dy -= cl->Ly * round(dy / cl->Ly);
                    double r_sq = dx * dx + dy * dy;
                    if (r_sq <= cl->rcutsq) W += calc_w(c, r_sq);
                }
            }
    }
    orc_cl_destroy(cl);
    return 0.5 * W;
}
/* :235-242 */
double orc_pressure_celllist(const orc_calc *c, double Lx, double Ly, const double *xy, int N) {
    double V = Lx * Ly;
    double rho = (double)N / V;
    double W = orc_total_virial_celllist(c, Lx, Ly, xy, N);
    return rho * c->T + W / (2.0 * V);
}

void orc_cell_ids(double Lx, double Ly, double rcut, const double *xy, int N, int *out_cell) {
    orc_celllist *cl = orc_cl_create(Lx, Ly, rcut);
    for (int i = 0; i < N; ++i) out_cell[i] = orc_cl_cell_index(cl, xy[2 * i], xy[2 * i + 1]);
    orc_cl_destroy(cl);
}
void orc_neighbor_counts(double Lx, double Ly, double rcut, const double *xy, int N, int *out_count) {
    orc_celllist *cl = orc_cl_create(Lx, Ly, rcut);
    orc_cl_build(cl, xy, N);
    for (int i = 0; i < N; ++i) out_count[i] = orc_cl_neighbors(cl, i, xy, NULL, NULL, 0);
    orc_cl_destroy(cl);
}
double orc_local_energy(const orc_calc *c, double Lx, double Ly, const double *xy, int N, int idx) {
    orc_celllist *cl = orc_cl_create(Lx, Ly, c->rcut);
    orc_cl_build(cl, xy, N);
    double u = scan_sum(cl, c, xy, idx, 0);
    orc_cl_destroy(cl);
    return u;
}
double orc_point_energy(const orc_calc *c, double Lx, double Ly, const double *xy, int N, double x, double y,
                        int exclude_idx) {
    orc_celllist *cl = orc_cl_create(Lx, Ly, c->rcut);
    orc_cl_build(cl, xy, N);
    double u = scan_sum_point(cl, c, xy, x, y, exclude_idx);
    orc_cl_destroy(cl);
    return u;
}

/* ------------------------------------------------------------------------- */
/* RNG — serial_src/rng.cpp over libstdc++ <random>                            */
/* ------------------------------------------------------------------------- */

/* std::mt19937 (MT19937, Matsumoto & Nishimura 1998), seeded as
 * std::mersenne_twister_engine::seed(value) does. */
void orc_rng_seed(orc_rng *r, uint32_t seed) {
    r->mt[0] = seed;
    for (int i = 1; i < 624; ++i) r->mt[i] = 1812433253u * (r->mt[i - 1] ^ (r->mt[i - 1] >> 30)) + (uint32_t)i;
    r->idx = 624;
}
uint32_t orc_rng_next_u32(orc_rng *r) {
    if (r->idx >= 624) {
        uint32_t *mt = r->mt;
        for (int k = 0; k < 624; ++k) {
            uint32_t y = (mt[k] & 0x80000000u) | (mt[(k + 1) % 624] & 0x7fffffffu);
            mt[k] = mt[(k + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
        }
        r->idx = 0;
    }
    uint32_t y = r->mt[r->idx++];
    y ^= (y >> 11);
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= (y >> 18);
    return y;
}
/* rng.cpp:9-11: uniform_real_distribution<double>(0,1) = libstdc++ generate_canonical<double,53>
 * over a 32-bit engine: two draws, low word first, (lo + hi*2^32) / 2^64, clamp below 1. */
double orc_rng_uniform01(orc_rng *r) {
    double sum = 0.0, tmp = 1.0;
    for (int k = 0; k < 2; ++k) {
        sum += (double)orc_rng_next_u32(r) * tmp;
        tmp *= 4294967296.0;
    }
    double ret = sum / tmp;
    if (ret >= 1.0) ret = nextafter(1.0, 0.0);
    return ret * (1.0 - 0.0) + 0.0;
}
/* rng.cpp:14-16 */
double orc_rng_uniform(orc_rng *r, double a, double b) { return a + (b - a) * orc_rng_uniform01(r); }
/* rng.cpp:19-22: a fresh uniform_int_distribution<int> per call; libstdc++ (GCC >= 11) downscales a
 * 32-bit engine with Lemire's nearly-divisionless method (bits/uniform_int_dist.h, _S_nd). */
int orc_rng_uniform_int(orc_rng *r, int lo, int hi) {
    uint64_t urange = (uint64_t)(uint32_t)hi - (uint64_t)(uint32_t)lo; /* __uctype arithmetic */
    urange &= 0xffffffffull;
    uint64_t ret;
    if (0xffffffffull > urange) {
        uint32_t range = (uint32_t)(urange + 1);
        uint64_t product = (uint64_t)orc_rng_next_u32(r) * (uint64_t)range;
        uint32_t low = (uint32_t)product;
        if (low < range) {
            uint32_t threshold = (uint32_t)(-range) % range;
            while (low < threshold) {
                product = (uint64_t)orc_rng_next_u32(r) * (uint64_t)range;
                low = (uint32_t)product;
            }
        }
        ret = product >> 32;
    } else {
        ret = orc_rng_next_u32(r); /* full 32-bit range */
    }
    return (int)(ret + (uint64_t)(uint32_t)lo);
}

/* initial.cpp:86-104, random branch */
int orc_initialize_particles_random(double Lx, double Ly, int N, unsigned seed, double *xy) {
    if (seed == 0) return -1; /* the reference seeds from time(0): not reproducible */
    srand(seed + 10);
    for (int i = 0; i < N; ++i) {
        double x = (double)rand() / RAND_MAX * Lx;
        double y = (double)rand() / RAND_MAX * Ly;
        xy[2 * i] = x;
        xy[2 * i + 1] = y;
    }
    return 0;
}

/* ------------------------------------------------------------------------- */
/* the serial step loop — serial_src/MC.cpp                                    */
/* ------------------------------------------------------------------------- */

struct orc_mc {
    double Lx, Ly, V, invLx, invLy;
    double *xy;
    int N, cap;
    orc_calc calc;
    orc_celllist *cl;
    orc_rng rng;
    double rcut, delta;
    int ensemble;
    double beta, z, energy;
    long long step_time; /* simulation_step_time, MC.h:85 */
    long long accepted, total;
};

static void mc_reserve(orc_mc *mc, int n) {
    if (n <= mc->cap) return;
    mc->cap = n + n / 2 + 64;
    mc->xy = (double *)realloc(mc->xy, sizeof(double) * 2 * (size_t)mc->cap);
}

/* MC.cpp:4-29 */
orc_mc *orc_mc_create(double Lx, double Ly, const double *xy, int N, const orc_calc *calc, double rcut,
                      double delta, int ensemble, uint32_t rng_seed) {
    orc_mc *mc = (orc_mc *)calloc(1, sizeof(*mc));
    mc->Lx = Lx;
    mc->Ly = Ly;
    mc->invLx = 1.0 / Lx;
    mc->invLy = 1.0 / Ly;
    mc->V = Lx * Ly;
    mc_reserve(mc, N);
    if (N > 0) memcpy(mc->xy, xy, sizeof(double) * 2 * (size_t)N);
    mc->N = N;
    mc->calc = *calc;
    mc->cl = orc_cl_create(Lx, Ly, rcut);
    orc_rng_seed(&mc->rng, rng_seed);
    mc->rcut = rcut;
    mc->delta = delta;
    mc->ensemble = ensemble;
    mc->beta = 1.0 / calc->T;
    mc->z = calc->z;
    mc->energy = orc_total_energy_celllist(calc, Lx, Ly, mc->xy, N);
    orc_cl_build(mc->cl, mc->xy, N);
    return mc;
}
void orc_mc_destroy(orc_mc *mc) {
    if (!mc) return;
    orc_cl_destroy(mc->cl);
    free(mc->xy);
    free(mc);
}
/* MC.cpp:251-253 */
void orc_mc_update_celllist(orc_mc *mc) { orc_cl_build(mc->cl, mc->xy, mc->N); }

static void mc_apply_pbc(const orc_mc *mc, double *p) {
    /* initial.cpp:22-25 with the box's precomputed reciprocals */
    p[0] -= mc->Lx * floor(p[0] * mc->invLx);
    p[1] -= mc->Ly * floor(p[1] * mc->invLy);
}

/* MC.cpp:96-123.  The neighbour scans use particle_cell_ from the LAST build (stale-cell quirk,
 * cell_list.cpp:45). */
int orc_mc_step_celllist(orc_mc *mc) {
    int idx = orc_rng_uniform_int(&mc->rng, 0, mc->N - 1);
    double oldx = mc->xy[2 * idx], oldy = mc->xy[2 * idx + 1];
    double dx = orc_rng_uniform(&mc->rng, -mc->delta, mc->delta);
    double dy = orc_rng_uniform(&mc->rng, -mc->delta, mc->delta);
    double U_old = scan_sum(mc->cl, &mc->calc, mc->xy, idx, 0);
    mc->xy[2 * idx] += dx;
    mc->xy[2 * idx + 1] += dy;
    mc_apply_pbc(mc, &mc->xy[2 * idx]);
    double U_new = scan_sum(mc->cl, &mc->calc, mc->xy, idx, 0);
    double dU = U_new - U_old;
    mc->energy += dU;
    int accept = 0;
    if (dU <= 0.0) {
        accept = 1;
    } else {
        double p = exp(-mc->beta * dU);
        accept = (orc_rng_uniform01(&mc->rng) < p);
    }
    if (!accept) {
        mc->xy[2 * idx] = oldx;
        mc->xy[2 * idx + 1] = oldy;
        mc->energy -= dU;
    }
    return accept;
}
/* MC.cpp:125-151 */
int orc_mc_step_bruteforce(orc_mc *mc) {
    int idx = orc_rng_uniform_int(&mc->rng, 0, mc->N - 1);
    double oldx = mc->xy[2 * idx], oldy = mc->xy[2 * idx + 1];
    double dx = orc_rng_uniform(&mc->rng, -mc->delta, mc->delta);
    double dy = orc_rng_uniform(&mc->rng, -mc->delta, mc->delta);
    double U_old = orc_total_energy(&mc->calc, mc->Lx, mc->Ly, mc->xy, mc->N);
    mc->xy[2 * idx] += dx;
    mc->xy[2 * idx + 1] += dy;
    mc_apply_pbc(mc, &mc->xy[2 * idx]);
    double U_new = orc_total_energy(&mc->calc, mc->Lx, mc->Ly, mc->xy, mc->N);
    double dU = U_new - U_old;
    mc->energy += dU;
    int accept = 0;
    if (dU <= 0.0) {
        accept = 1;
    } else {
        double p = exp(-mc->beta * dU);
        accept = (orc_rng_uniform01(&mc->rng) < p);
    }
    if (!accept) {
        mc->xy[2 * idx] = oldx;
        mc->xy[2 * idx + 1] = oldy;
        mc->energy -= dU;
    }
    return accept;
}
/* MC.cpp:189-249 */
int orc_mc_gc_move(orc_mc *mc) {
    double V = mc->V;
    size_t N = (size_t)mc->N;
    double check = orc_rng_uniform01(&mc->rng);
    if (check < 0.3) {
        /* MC.cpp:198 draws both coordinates inside one constructor call; g++ evaluates the
         * arguments right to left, so y is drawn BEFORE x (pinned against oracle/_ref). */
        double py = orc_rng_uniform(&mc->rng, 0, mc->Ly);
        double px = orc_rng_uniform(&mc->rng, 0, mc->Lx);
        double dU = scan_sum_point(mc->cl, &mc->calc, mc->xy, px, py, -1);
        double acc = (mc->z * V) / (N + 1) * exp(-mc->beta * dU);
        int accept = 0;
        if (acc >= 1) accept = 1;
        else if (orc_rng_uniform01(&mc->rng) < acc) accept = 1;
        if (accept) {
            mc->energy += dU;
            mc_reserve(mc, mc->N + 1);
            mc->xy[2 * mc->N] = px;
            mc->xy[2 * mc->N + 1] = py;
            mc->N += 1;
            orc_mc_update_celllist(mc);
            return 1;
        }
        return 0;
    } else if (check < 0.6) {
        if (N == 0) return 0;
        int idx = orc_rng_uniform_int(&mc->rng, 0, (int)(N - 1));
        double dU = -scan_sum(mc->cl, &mc->calc, mc->xy, idx, 0);
        double acc = N / (mc->z * V) * exp(-mc->beta * dU);
        int accept = 0;
        if (acc >= 1) accept = 1;
        else if (orc_rng_uniform01(&mc->rng) < acc) accept = 1;
        if (accept) {
            memmove(&mc->xy[2 * idx], &mc->xy[2 * (idx + 1)], sizeof(double) * 2 * (size_t)(mc->N - idx - 1));
            mc->N -= 1;
            orc_mc_update_celllist(mc);
            mc->energy += dU;
            return 1;
        }
        return 0;
    }
    return 0;
}
/* MC.cpp:31-78 (NVT and GCMC branches; NPT is outside the hot path) */
int orc_mc_run(orc_mc *mc, size_t nSteps, size_t fOutputStep, size_t fUpdateCell, long long *s_time, int *s_N,
               double *s_E, int cap) {
    int ns = 0;
    for (size_t step = 0; step < nSteps; ++step) {
        size_t N = (size_t)mc->N;
        for (size_t i = 0; i < N; ++i) {
            int a = orc_mc_step_celllist(mc);
            mc->step_time += 1;
            if (a) mc->accepted++;
            mc->total += 1;
        }
        if (mc->ensemble == ORC_GCMC) {
            int a = orc_mc_gc_move(mc);
            mc->step_time += 1;
            if (a) mc->accepted++;
            mc->total += 1;
        }
        if (step % fUpdateCell == 0) {
            orc_mc_update_celllist(mc);
            mc->energy = orc_total_energy_celllist(&mc->calc, mc->Lx, mc->Ly, mc->xy, mc->N);
        }
        if ((step + 1) % fOutputStep == 0 || step == nSteps - 1) {
            if (ns < cap) {
                if (s_time) s_time[ns] = mc->step_time;
                if (s_N) s_N[ns] = mc->N;
                if (s_E) s_E[ns] = mc->energy;
            }
            ++ns;
        }
    }
    return ns;
}
double orc_mc_energy(const orc_mc *mc) { return mc->energy; }
int orc_mc_n(const orc_mc *mc) { return mc->N; }
void orc_mc_get_particles(const orc_mc *mc, double *xy) {
    if (mc->N > 0) memcpy(xy, mc->xy, sizeof(double) * 2 * (size_t)mc->N);
}
long long orc_mc_accepted(const orc_mc *mc) { return mc->accepted; }
long long orc_mc_total(const orc_mc *mc) { return mc->total; }

/* ------------------------------------------------------------------------- */
/* golden-input generator (sanity_check frames)                                */
/* ------------------------------------------------------------------------- */

/* std::seed_seq::generate ([rand.util.seedseq]) for the 5-word sequence of rng_parallel.h:18-24 */
static void seed_seq_generate(const uint32_t *v, int s, uint32_t *b, int n) {
    for (int i = 0; i < n; ++i) b[i] = 0x8b8b8b8bu;
    int t = (n >= 623) ? 11 : (n >= 68) ? 7 : (n >= 39) ? 5 : (n >= 7) ? 3 : (n - 1) / 2;
    int p = (n - t) / 2;
    int q = p + t;
    int m = (s + 1 > n) ? s + 1 : n;
    for (int k = 0; k < m; ++k) {
        uint32_t arg = b[k % n] ^ b[(k + p) % n] ^ b[(k + n - 1) % n];
        uint32_t r1 = 1664525u * (arg ^ (arg >> 27));
        uint32_t r2 = r1;
        if (k == 0) r2 += (uint32_t)s;
        else if (k <= s) r2 += (uint32_t)(k % n) + v[k - 1];
        else r2 += (uint32_t)(k % n);
        b[(k + p) % n] += r1;
        b[(k + q) % n] += r2;
        b[k % n] = r2;
    }
    for (int k = m; k < m + n; ++k) {
        uint32_t arg = b[k % n] + b[(k + p) % n] + b[(k + n - 1) % n];
        uint32_t r3 = 1566083941u * (arg ^ (arg >> 27));
        uint32_t r4 = r3 - (uint32_t)(k % n);
        b[(k + p) % n] ^= r3;
        b[(k + q) % n] ^= r4;
        b[k % n] = r4;
    }
}
/* std::mt19937_64 */
typedef struct {
    uint64_t mt[312];
    int idx;
} mt64;
static void mt64_seed_seq(mt64 *g, const uint32_t *v, int s) {
    uint32_t arr[624];
    seed_seq_generate(v, s, arr, 624);
    int zero = 1;
    for (int i = 0; i < 312; ++i) {
        g->mt[i] = (uint64_t)arr[2 * i] + ((uint64_t)arr[2 * i + 1] << 32);
        if (zero) {
            if (i == 0) {
                if ((g->mt[0] & (~0ull << 31)) != 0) zero = 0;
            } else if (g->mt[i] != 0)
                zero = 0;
        }
    }
    if (zero) g->mt[0] = 1ull << 63;
    g->idx = 312;
}
static uint64_t mt64_next(mt64 *g) {
    if (g->idx >= 312) {
        const uint64_t UM = ~0ull << 31, LM = ~UM;
        for (int k = 0; k < 312; ++k) {
            uint64_t y = (g->mt[k] & UM) | (g->mt[(k + 1) % 312] & LM);
            g->mt[k] = g->mt[(k + 156) % 312] ^ (y >> 1) ^ ((y & 1ull) ? 0xb5026f5aa96619e9ull : 0ull);
        }
        g->idx = 0;
    }
    uint64_t z = g->mt[g->idx++];
    z ^= (z >> 29) & 0x5555555555555555ull;
    z ^= (z << 17) & 0x71d67fffeda60000ull;
    z ^= (z << 37) & 0xfff7eee000000000ull;
    z ^= (z >> 43);
    return z;
}
/* generate_canonical<double,53> over a 64-bit engine: one draw / 2^64, clamped below 1 */
static double mt64_uniform01(mt64 *g) {
    double ret = (double)mt64_next(g) / 18446744073709551616.0;
    if (ret >= 1.0) ret = nextafter(1.0, 0.0);
    return ret;
}

/* mpi_src/simulatin_box.cpp:100-128 */
void orc_best_decomposition(double Lx, double Ly, int nranks, int *Px, int *Py) {
    int bestPx = 1, bestPy = nranks;
    double best = INFINITY;
    for (int px = 1; px * px <= nranks; ++px) {
        if (nranks % px) continue;
        int py = nranks / px;
        double s1 = fabs(log((Lx / px) / (Ly / py))) + abs(px - py) * 1e-3;
        if (s1 < best) { best = s1; bestPx = px; bestPy = py; }
        double s2 = fabs(log((Lx / py) / (Ly / px))) + abs(py - px) * 1e-3;
        if (s2 < best) { best = s2; bestPx = py; bestPy = px; }
    }
    *Px = bestPx;
    *Py = bestPy;
}

int orc_sanity_frame(double Lx, double Ly, int Px, int Py, int N_global, unsigned seed, double *xy) {
    int P = Px * Py, out = 0;
    for (int rank = 0; rank < P; ++rank) {
        /* initial_mpi.cpp:5-10 */
        int base = N_global / P, rem = N_global % P;
        int n_local = base + (rank < rem ? 1 : 0);
        /* simulatin_box.cpp:91-98 */
        int cx = rank % Px, cy = rank / Px;
        double ddx = Lx / Px, ddy = Ly / Py;
        double x0 = cx * ddx, x1 = (cx + 1) * ddx, y0 = cy * ddy, y1 = (cy + 1) * ddy;
        double dx = x1 - x0, dy = y1 - y0;
        /* rng_parallel.h:16-27 */
        uint32_t v[5] = {seed, (uint32_t)rank, 0x9E3779B9u, 0x85EBCA6Bu, 0xC2B2AE35u};
        mt64 g;
        mt64_seed_seq(&g, v, 5);
        /* initial_mpi.cpp:41-51 */
        for (int i = 0; i < n_local; ++i) {
            double xr = mt64_uniform01(&g);
            double yr = mt64_uniform01(&g);
            xy[2 * out] = x0 + xr * dx;
            xy[2 * out + 1] = y0 + yr * dy;
            ++out;
        }
    }
    return out;
}
