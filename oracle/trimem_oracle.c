/* TEST INFRASTRUCTURE ONLY — the product never links, imports or executes this file.
 *
 * Plain-C restatement of the reference's CPU algorithm for the force-evaluation hot
 * path (bio-phys/trimem @ 1a2a05c), written function by function after the reference
 * sources and deliberately in the reference's own formulation (acos / tan of sector
 * angles, per-half-edge recomputation, repeated additions for multiplicities), i.e.
 * NOT in the algebraic form the CUDA kernels use. It works on flattened half-edge
 * arrays (he_to, he_next, he_face, v_he; opposite(h) = h^1) that oracle/halfedge.py
 * builds with OpenMesh's add_face numbering.
 *
 * Parity status: PINNED. tests/test_oracle.py checks this file against the committed
 * golden vectors generated from the unmodified reference sources (oracle/_ref, see
 * tests/golden/make_golden.py) and against the docs' known answer
 * 1000755.3602085959 (reference doc/python-module.rst:88-97).
 *
 * Third-party arithmetic restated here: OpenMesh (pinned db8c0608..., absent from the
 * reference tree) calc_edge_vector / calc_sector_normal / calc_sector_angle /
 * calc_dihedral_angle, per OpenMesh's published PolyMeshT implementation.
 */
#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    int V, HE;
    const int *he_to, *he_next, *he_face, *v_he;
    const double* x; /* V*3 */
} to_mesh;

typedef struct {
    double kappa_b, kappa_a, kappa_v, kappa_c, kappa_t, kappa_r;
    double ref_area, ref_volume, ref_curvature;
    double bond_lc0, bond_lc1, bond_a0;
    int bond_r, bond_type; /* 0 edge, 1 area */
    double rep_lc1;
    int rep_r, pad;
} to_params;

/* neighbour list: kind 0 = cell list, 1 = verlet list */
typedef struct {
    int kind, excl, V;
    double rlist;
    double box_min[3], box_max[3], r_list[3];
    int shape[3], strides[3];
    long ncell;
    int* cell_start; /* ncell+1 */
    int* cell_items; /* V, ascending id inside a cell (insertion order of the reference) */
    int* row_off;    /* verlet: V+1 */
    int* row_j;
    long n_entries;
} to_nlist;

/* ---------------------------------------------------------------- vectors */
typedef struct { double v[3]; } vec3;
static vec3 vsub(vec3 a, vec3 b) { vec3 r = {{a.v[0] - b.v[0], a.v[1] - b.v[1], a.v[2] - b.v[2]}}; return r; }
static vec3 vadd(vec3 a, vec3 b) { vec3 r = {{a.v[0] + b.v[0], a.v[1] + b.v[1], a.v[2] + b.v[2]}}; return r; }
static vec3 vscale(vec3 a, double s) { vec3 r = {{a.v[0] * s, a.v[1] * s, a.v[2] * s}}; return r; }
static vec3 vdiv(vec3 a, double s) { vec3 r = {{a.v[0] / s, a.v[1] / s, a.v[2] / s}}; return r; }
static double vdot(vec3 a, vec3 b) { double r = 0; for (int i = 0; i < 3; i++) r += a.v[i] * b.v[i]; return r; }
static double vnorm(vec3 a) { return sqrt(vdot(a, a)); }
static vec3 vcross(vec3 a, vec3 b) {
    vec3 r = {{a.v[1] * b.v[2] - a.v[2] * b.v[1], a.v[2] * b.v[0] - a.v[0] * b.v[2], a.v[0] * b.v[1] - a.v[1] * b.v[0]}};
    return r;
}
static vec3 point(const to_mesh* m, int v) { vec3 r = {{m->x[3 * v], m->x[3 * v + 1], m->x[3 * v + 2]}}; return r; }

/* ------------------------------------------- OpenMesh geometry, restated */
static int he_prev(const to_mesh* m, int h) {
    /* triangle faces: prev = next(next); boundary loops: walk */
    if (m->he_face[h] >= 0) return m->he_next[m->he_next[h]];
    int p = h;
    while (m->he_next[p] != h) p = m->he_next[p];
    return p;
}
static int from_v(const to_mesh* m, int h) { return m->he_to[h ^ 1]; }
static vec3 calc_edge_vector(const to_mesh* m, int h) { return vsub(point(m, m->he_to[h]), point(m, from_v(m, h))); }
static void calc_sector_vectors(const to_mesh* m, int in, vec3* v0, vec3* v1) {
    *v0 = calc_edge_vector(m, m->he_next[in]);
    *v1 = calc_edge_vector(m, in ^ 1);
}
static vec3 calc_sector_normal(const to_mesh* m, int in) {
    vec3 v0, v1;
    calc_sector_vectors(m, in, &v0, &v1);
    return vcross(v0, v1);
}
static double sane_aarg(double a) { if (a < -1) a = -1; else if (a > 1) a = 1; return a; }
static double angle_cs(double c, double s) { c = sane_aarg(c); return s >= 0 ? acos(c) : -acos(c); }
static double calc_sector_angle(const to_mesh* m, int in) {
    vec3 v0, v1;
    calc_sector_vectors(m, in, &v0, &v1);
    double denom = vnorm(v0) * vnorm(v1);
    if (denom == 0.0) return 0;
    double cos_a = vdot(v0, v1) / denom;
    if (m->he_face[in] < 0) {
        vec3 fn = calc_sector_normal(m, in ^ 1);
        double sign_a = vdot(vcross(v0, v1), fn);
        return angle_cs(cos_a, sign_a);
    }
    return acos(sane_aarg(cos_a));
}
static double calc_dihedral_angle(const to_mesh* m, int h) {
    if (m->he_face[h] < 0 || m->he_face[h ^ 1] < 0) return 0;
    vec3 n0 = calc_sector_normal(m, h), n1 = calc_sector_normal(m, h ^ 1), he = calc_edge_vector(m, h);
    double denom = vnorm(n0) * vnorm(n1);
    if (denom == 0.0) return 0;
    double da_cos = vdot(n0, n1) / denom;
    double da_sin_sign = vdot(vcross(n0, n1), he);
    return angle_cs(da_cos, da_sin_sign);
}

/* -------------------------------- mesh_util.h:15-185 (primitives + grads) */
static vec3 face_normal(const to_mesh* m, int h) { vec3 n = calc_sector_normal(m, h); return vdiv(n, vnorm(n)); }
static double edge_length(const to_mesh* m, int h) { return vnorm(calc_edge_vector(m, h)); }
static double face_area(const to_mesh* m, int h) { return vnorm(calc_sector_normal(m, h)) / 2; }
static double face_volume(const to_mesh* m, int h) {
    vec3 p0 = point(m, from_v(m, h)), p1 = point(m, m->he_to[h]), p2 = point(m, m->he_to[m->he_next[h]]);
    return vdot(vcross(p1, p2), p0) / 6;
}
/* gradient wrt the from-vertex only (template arg <1> in the reference) */
static vec3 edge_length_grad0(const to_mesh* m, int h) {
    vec3 e = calc_edge_vector(m, h);
    e = vdiv(e, vnorm(e));
    return vscale(e, -1.0);
}
static vec3 face_area_grad0(const to_mesh* m, int h) {
    vec3 n = face_normal(m, h);
    n = vdiv(n, vnorm(n));
    vec3 e1 = calc_edge_vector(m, m->he_next[h]);
    return vscale(vcross(n, e1), 0.5);
}
static vec3 face_volume_grad0(const to_mesh* m, int h) {
    vec3 p1 = point(m, m->he_to[h]), p2 = point(m, m->he_to[m->he_next[h]]);
    return vdiv(vcross(p1, p2), 6);
}
/* dihedral_angle_grad: bit 0 (from-vertex) and bit 2 (apex of the face of h) */
static vec3 dihedral_angle_grad_bit0(const to_mesh* m, int h) {
    vec3 n0 = face_normal(m, h), n1 = face_normal(m, h ^ 1);
    double l = edge_length(m, h);
    double a03 = calc_sector_angle(m, h);
    double a04 = calc_sector_angle(m, he_prev(m, h ^ 1));
    return vdiv(vadd(vscale(n0, 1. / tan(a03)), vscale(n1, 1. / tan(a04))), l);
}
static vec3 dihedral_angle_grad_bit2(const to_mesh* m, int h) {
    vec3 n0 = face_normal(m, h);
    double l = edge_length(m, h);
    double a01 = calc_sector_angle(m, he_prev(m, h));
    double a03 = calc_sector_angle(m, h);
    return vdiv(vscale(n0, -(1. / tan(a01) + 1. / tan(a03))), l);
}
/* all rows, for the single half-edge API (bindings.cpp:249-407) */
void to_halfedge_geometry(const to_mesh* m, int h, double* out /*40*/) {
    int i = from_v(m, h), j = m->he_to[h], k = m->he_to[m->he_next[h]];
    (void)i; (void)j; (void)k;
    out[0] = edge_length(m, h); out[1] = face_area(m, h); out[2] = face_volume(m, h); out[3] = calc_dihedral_angle(m, h);
    vec3 e = calc_edge_vector(m, h); e = vdiv(e, vnorm(e));
    for (int c = 0; c < 3; c++) { out[4 + c] = -e.v[c]; out[7 + c] = e.v[c]; }
    vec3 n = face_normal(m, h); n = vdiv(n, vnorm(n));
    vec3 g0 = vscale(vcross(n, calc_edge_vector(m, m->he_next[h])), 0.5);
    vec3 g1 = vscale(vcross(n, calc_edge_vector(m, he_prev(m, h))), 0.5);
    vec3 g2 = vscale(vcross(n, calc_edge_vector(m, h)), 0.5);
    vec3 p0 = point(m, from_v(m, h)), p1 = point(m, m->he_to[h]), p2 = point(m, m->he_to[m->he_next[h]]);
    vec3 v0 = vdiv(vcross(p1, p2), 6), v1 = vdiv(vcross(p2, p0), 6), v2 = vdiv(vcross(p0, p1), 6);
    vec3 n0 = face_normal(m, h), n1 = face_normal(m, h ^ 1);
    double l = edge_length(m, h);
    double a01 = calc_sector_angle(m, he_prev(m, h)), a03 = calc_sector_angle(m, h);
    double a02 = calc_sector_angle(m, h ^ 1), a04 = calc_sector_angle(m, he_prev(m, h ^ 1));
    vec3 d0 = vdiv(vadd(vscale(n0, 1. / tan(a03)), vscale(n1, 1. / tan(a04))), l);
    vec3 d1 = vdiv(vadd(vscale(n0, 1. / tan(a01)), vscale(n1, 1. / tan(a02))), l);
    vec3 d2 = vdiv(vscale(n0, -(1. / tan(a01) + 1. / tan(a03))), l);
    vec3 d3 = vdiv(vscale(n1, -(1. / tan(a02) + 1. / tan(a04))), l);
    vec3 rows[10] = {g0, g1, g2, v0, v1, v2, d0, d1, d2, d3};
    for (int r = 0; r < 10; r++) for (int c = 0; c < 3; c++) out[10 + 3 * r + c] = rows[r].v[c];
}

/* ------------------------------------------ mesh_tether.h:29-121 (bonds) */
static double bond_property(const to_mesh* m, const to_params* p, int h) {
    if (p->bond_type == 0) {
        double l = edge_length(m, h);
        if (l > p->bond_lc0) return pow(p->bond_r, p->bond_r + 1) * pow(l - p->bond_lc0, p->bond_r);
        if (l < p->bond_lc1) return exp(l / (l - p->bond_lc1)) / pow(l, p->bond_r);
        return 0.0;
    }
    double a = face_area(m, h), d = a / p->bond_a0 - 1.0;
    return d * d;
}
static vec3 bond_property_grad(const to_mesh* m, const to_params* p, int h) {
    if (p->bond_type == 0) {
        double l = edge_length(m, h);
        vec3 lg = edge_length_grad0(m, h);
        double fac = 0.0;
        if (l > p->bond_lc0) fac = pow(p->bond_r, p->bond_r + 2) * pow(l - p->bond_lc0, p->bond_r - 1);
        if (l < p->bond_lc1) {
            double lr = pow(l, p->bond_r), lmc = l - p->bond_lc1;
            fac -= p->bond_lc1 / (lmc * lmc) + p->bond_r / l;
            fac /= lr;
            fac *= exp(l / lmc);
        }
        return vscale(lg, fac);
    }
    double a = face_area(m, h);
    vec3 ag = face_area_grad0(m, h);
    double d = 2.0 * (a / p->bond_a0 - 1.0) / p->bond_a0;
    return vscale(ag, d);
}
static int bond_valence(const to_params* p) { return p->bond_type == 0 ? 2 : 3; }

/* -------------------------------------------- nlists: cell_list.h:43-129 */
static int in_exclusion(const to_mesh* m, int pid, int jid, int excl) {
    /* cell_list.h:167-201: walk the ring (and ring of ring) of pid on the live mesh */
    if (excl <= 0) return 0;
    int h0 = m->v_he[pid];
    if (h0 < 0) return 0;
    int h = h0;
    do {
        int vii = m->he_to[h];
        if (vii == jid) return 1;
        if (excl > 1) {
            int g0 = m->v_he[vii], g = g0;
            if (g0 >= 0) do { if (m->he_to[g] == jid) return 1; g = m->he_next[g ^ 1]; } while (g != g0);
        }
        h = m->he_next[h ^ 1];
    } while (h != h0);
    return 0;
}

static double distance3(const double* a, const double* b) {
    double dist = 0.0;
    for (int i = 0; i < 3; i++) { double di = a[i] - b[i]; dist += di * di; }
    return sqrt(dist);
}

void to_nlist_free(to_nlist* nl) {
    if (!nl) return;
    free(nl->cell_start); free(nl->cell_items); free(nl->row_off); free(nl->row_j);
    free(nl);
}

static int cell_id(const to_nlist* nl, const double* pt) {
    int id = 0;
    for (int k = 0; k < 3; k++) id += (int)((pt[k] - nl->box_min[k]) / nl->r_list[k]) * nl->strides[k];
    return id;
}

static to_nlist* cell_list_build(const to_mesh* m, double rlist, int excl, double box_eps) {
    to_nlist* nl = (to_nlist*)calloc(1, sizeof(to_nlist));
    nl->kind = 0; nl->excl = excl; nl->V = m->V; nl->rlist = rlist;
    /* `box_min = { DBL_MAX }`, `box_max = { -1 }` brace-initialise only the first component */
    nl->box_min[0] = DBL_MAX; nl->box_min[1] = 0; nl->box_min[2] = 0;
    nl->box_max[0] = -1; nl->box_max[1] = 0; nl->box_max[2] = 0;
    for (int i = 0; i < m->V; i++)
        for (int k = 0; k < 3; k++) {
            double v = m->x[3 * i + k];
            if (v < nl->box_min[k]) nl->box_min[k] = v;
            if (v > nl->box_max[k]) nl->box_max[k] = v;
        }
    nl->ncell = 1;
    for (int k = 0; k < 3; k++) {
        double dim = nl->box_max[k] - nl->box_min[k] + box_eps;
        int nk = (int)(dim / rlist);
        nl->r_list[k] = (nk > 0) ? dim / nk : dim;
        nl->shape[k] = (nk > 0) ? nk : 1;
        nl->ncell *= nl->shape[k];
    }
    nl->strides[0] = 1; nl->strides[1] = nl->shape[0]; nl->strides[2] = nl->shape[0] * nl->shape[1];
    /* std::map<int, vector<int>> cells -> dense counting sort (same content, ids ascending) */
    nl->cell_start = (int*)calloc((size_t)nl->ncell + 1, sizeof(int));
    nl->cell_items = (int*)malloc(sizeof(int) * (size_t)(m->V > 0 ? m->V : 1));
    int* cid = (int*)malloc(sizeof(int) * (size_t)(m->V > 0 ? m->V : 1));
    for (int i = 0; i < m->V; i++) { cid[i] = cell_id(nl, m->x + 3 * i); nl->cell_start[cid[i] + 1]++; }
    for (long c = 0; c < nl->ncell; c++) nl->cell_start[c + 1] += nl->cell_start[c];
    int* cur = (int*)malloc(sizeof(int) * (size_t)nl->ncell);
    memcpy(cur, nl->cell_start, sizeof(int) * (size_t)nl->ncell);
    for (int i = 0; i < m->V; i++) nl->cell_items[cur[cid[i]]++] = i;
    free(cur); free(cid);
    return nl;
}

/* visit candidates of pid: cell_list.h:131-219 / verlet_list.h:74-101 */
typedef void (*pair_cb)(void* ctx, int j, vec3 d);
static void point_distances_cb(const to_nlist* nl, const to_mesh* m, int pid, double dmax, pair_cb cb, void* ctx) {
    vec3 pt = point(m, pid);
    if (nl->kind == 1) {
        for (int k = nl->row_off[pid]; k < nl->row_off[pid + 1]; k++) {
            int jid = nl->row_j[k];
            vec3 di = vsub(pt, point(m, jid));
            if (vnorm(di) < dmax) cb(ctx, jid, di);
        }
        return;
    }
    int coords[3];
    for (int k = 0; k < 3; k++) coords[k] = (int)((pt.v[k] - nl->box_min[k]) / nl->r_list[k]);
    for (int i = -1; i < 2; i++) {
        int ii = coords[0] + i;
        if (ii < 0 || ii >= nl->shape[0]) continue;
        for (int j = -1; j < 2; j++) {
            int jj = coords[1] + j;
            if (jj < 0 || jj >= nl->shape[1]) continue;
            for (int k = -1; k < 2; k++) {
                int kk = coords[2] + k;
                if (kk < 0 || kk >= nl->shape[2]) continue;
                int ocell = ii * nl->strides[0] + jj * nl->strides[1] + kk * nl->strides[2];
                for (int q = nl->cell_start[ocell]; q < nl->cell_start[ocell + 1]; q++) {
                    int jid = nl->cell_items[q];
                    if (pid == jid) continue;
                    if (in_exclusion(m, pid, jid, nl->excl)) continue;
                    vec3 dist = vsub(pt, point(m, jid));
                    if (vnorm(dist) < dmax) cb(ctx, jid, dist);
                }
            }
        }
    }
}

typedef struct { long n, cap; double* d; int* i; int* j; } coo;
static void coo_push(coo* c, double d, int i, int j) {
    if (c->n == c->cap) {
        c->cap = c->cap ? 2 * c->cap : 1024;
        c->d = (double*)realloc(c->d, sizeof(double) * c->cap);
        c->i = (int*)realloc(c->i, sizeof(int) * c->cap);
        c->j = (int*)realloc(c->j, sizeof(int) * c->cap);
    }
    c->d[c->n] = d; c->i[c->n] = i; c->j[c->n] = j; c->n++;
}

/* CellList::distance_matrix (cell_list.h:221-312): pairs i<j, dist <= rlist */
static void cell_distance_matrix(const to_nlist* nl, const to_mesh* m, double rlist, coo* out) {
    for (long icell = 0; icell < nl->ncell; icell++) {
        if (nl->cell_start[icell] == nl->cell_start[icell + 1]) continue;
        int kic = (int)(icell / nl->strides[2]), rem = (int)(icell % nl->strides[2]);
        int jic = rem / nl->strides[1], iic = rem % nl->strides[1];
        for (int a = -1; a < 2; a++) {
            int ii = iic + a; if (ii < 0 || ii >= nl->shape[0]) continue;
            for (int b = -1; b < 2; b++) {
                int jj = jic + b; if (jj < 0 || jj >= nl->shape[1]) continue;
                for (int c = -1; c < 2; c++) {
                    int kk = kic + c; if (kk < 0 || kk >= nl->shape[2]) continue;
                    long ocell = ii * nl->strides[0] + jj * nl->strides[1] + (long)kk * nl->strides[2];
                    if (nl->cell_start[ocell] == nl->cell_start[ocell + 1]) continue;
                    if (icell > ocell) continue;                 /* compute_cell_pairs: icell <= ocell */
                    int same = (icell == ocell);
                    for (int p = nl->cell_start[icell]; p < nl->cell_start[icell + 1]; p++) {
                        int iv = nl->cell_items[p];
                        for (int q = nl->cell_start[ocell]; q < nl->cell_start[ocell + 1]; q++) {
                            int ov = nl->cell_items[q];
                            if (iv == ov) continue;
                            if (in_exclusion(m, iv, ov, nl->excl)) continue;
                            if (same && iv > ov) continue;
                            double dist = distance3(m->x + 3 * iv, m->x + 3 * ov);
                            if (dist <= rlist) coo_push(out, dist, iv < ov ? iv : ov, iv < ov ? ov : iv);
                        }
                    }
                }
            }
        }
    }
}

to_nlist* to_nlist_build(const to_mesh* m, int kind, double rlist, int excl, double box_eps) {
    to_nlist* cl = cell_list_build(m, rlist, excl, box_eps);
    if (kind == 0) return cl;
    /* VerletList ctor (verlet_list.h:18-35): symmetrised distance_matrix(rlist) of a cell list */
    coo c = {0, 0, NULL, NULL, NULL};
    cell_distance_matrix(cl, m, rlist, &c);
    to_nlist* nl = (to_nlist*)calloc(1, sizeof(to_nlist));
    nl->kind = 1; nl->excl = excl; nl->V = m->V; nl->rlist = rlist;
    for (int k = 0; k < 3; k++) { nl->shape[k] = cl->shape[k]; nl->box_min[k] = cl->box_min[k]; nl->box_max[k] = cl->box_max[k]; nl->r_list[k] = cl->r_list[k]; }
    nl->row_off = (int*)calloc((size_t)m->V + 1, sizeof(int));
    for (long q = 0; q < c.n; q++) { nl->row_off[c.i[q] + 1]++; nl->row_off[c.j[q] + 1]++; }
    for (int v = 0; v < m->V; v++) nl->row_off[v + 1] += nl->row_off[v];
    nl->n_entries = nl->row_off[m->V];
    nl->row_j = (int*)malloc(sizeof(int) * (size_t)(nl->n_entries > 0 ? nl->n_entries : 1));
    int* cur = (int*)malloc(sizeof(int) * (size_t)(m->V > 0 ? m->V : 1));
    memcpy(cur, nl->row_off, sizeof(int) * (size_t)m->V);
    for (long q = 0; q < c.n; q++) { nl->row_j[cur[c.i[q]]++] = c.j[q]; nl->row_j[cur[c.j[q]]++] = c.i[q]; }
    free(cur); free(c.d); free(c.i); free(c.j);
    to_nlist_free(cl);
    return nl;
}

long to_nlist_entries(const to_nlist* nl) { return nl->n_entries; }
void to_nlist_grid(const to_nlist* nl, int* shape, double* box, double* cell) {
    for (int k = 0; k < 3; k++) { shape[k] = nl->shape[k]; box[k] = nl->box_min[k]; box[3 + k] = nl->box_max[k]; cell[k] = nl->r_list[k]; }
}

/* distance_matrix into malloc'ed arrays; returns the count */
long to_nlist_distance_matrix(const to_nlist* nl, const to_mesh* m, double rdist, double** d, int** i, int** j) {
    coo c = {0, 0, NULL, NULL, NULL};
    if (nl->kind == 0) cell_distance_matrix(nl, m, rdist, &c);
    else
        for (int v = 0; v < m->V; v++)          /* verlet_list.h:37-72: both directions, strict < */
            for (int k = nl->row_off[v]; k < nl->row_off[v + 1]; k++) {
                double dist = distance3(m->x + 3 * v, m->x + 3 * nl->row_j[k]);
                if (dist < rdist) coo_push(&c, dist, v, nl->row_j[k]);
            }
    *d = c.d; *i = c.i; *j = c.j;
    return c.n;
}
void to_free(void* p) { free(p); }

typedef struct { int n, cap; double* d; int* j; } pd_ctx;
static void pd_cb(void* ctx, int j, vec3 d) {
    pd_ctx* c = (pd_ctx*)ctx;
    if (c->n < c->cap) { for (int k = 0; k < 3; k++) c->d[3 * c->n + k] = d.v[k]; c->j[c->n] = j; }
    c->n++;
}
int to_nlist_point_distances(const to_nlist* nl, const to_mesh* m, int pid, double dmax, int cap, double* d, int* j) {
    pd_ctx c = {0, cap, d, j};
    point_distances_cb(nl, m, pid, dmax, pd_cb, &c);
    return c.n;
}

/* ----------------------------- mesh_repulsion.h:28-79 (flat-bottom repulsion) */
typedef struct { const to_params* p; double pot; vec3 grad; } rep_ctx;
static void rep_pot_cb(void* ctx, int j, vec3 d) {
    rep_ctx* c = (rep_ctx*)ctx; (void)j;
    double l = vnorm(d);
    c->pot += exp(l / (l - c->p->rep_lc1)) / pow(l, c->p->rep_r);
}
static void rep_grad_cb(void* ctx, int j, vec3 d) {
    rep_ctx* c = (rep_ctx*)ctx; (void)j;
    double fac = 0.0, l = vnorm(d), lr = pow(l, c->p->rep_r + 1), lmc = l - c->p->rep_lc1;
    fac -= c->p->rep_lc1 / (lmc * lmc) + c->p->rep_r / l;
    fac /= lr;
    fac *= exp(l / lmc);
    d = vscale(d, fac);
    c->grad = vadd(c->grad, vscale(d, 2));     /* `d_props[idx].repulsion += 2 * d_repulse[i]` */
}

/* ---------------------------- mesh_properties.cpp:12-49 (vertex_properties) */
void to_vertex_properties_one(const to_mesh* m, const to_params* p, const to_nlist* nl, int ve, double* out) {
    double area = 0, volume = 0, curvature = 0, tethering = 0;
    int h0 = m->v_he[ve];
    if (h0 >= 0) {
        int h = h0;
        do {
            if (m->he_face[h] >= 0) {
                double el = edge_length(m, h), da = calc_dihedral_angle(m, h);
                curvature += 0.5 * el * da;
                area += face_area(m, h);
                volume += face_volume(m, h);
                tethering += bond_property(m, p, h);
            }
            h = m->he_next[h ^ 1];
        } while (h != h0);
    }
    area /= 3; volume /= 3; curvature /= 2; tethering /= bond_valence(p);
    double bending = 2 * curvature * curvature / area;
    rep_ctx rc = {p, 0.0, {{0, 0, 0}}};
    if (nl) point_distances_cb(nl, m, ve, p->rep_lc1, rep_pot_cb, &rc);
    out[0] = area; out[1] = volume; out[2] = curvature; out[3] = bending; out[4] = tethering; out[5] = rc.pot;
}

void to_vertex_properties(const to_mesh* m, const to_params* p, const to_nlist* nl, double* vprops /*V*6*/) {
#pragma omp parallel for schedule(static)
    for (int i = 0; i < m->V; i++) to_vertex_properties_one(m, p, nl, i, vprops + 6 * i);
}

/* kernel.h:170-182,243-253: reduction; serial left-to-right here (the reference's OpenMP
 * combine order is unspecified) */
void to_reduce(int V, const double* vprops, double* out /*6*/) {
    for (int q = 0; q < 6; q++) out[q] = 0.0;
    for (int i = 0; i < V; i++) for (int q = 0; q < 6; q++) out[q] += vprops[6 * i + q];
}

/* kernel.h:110-122 */
double to_energy(const to_params* p, const double* s) {
    double e = 0.0, d;
    d = s[0] / p->ref_area - 1.0; e += p->kappa_a * d * d;
    d = s[1] / p->ref_volume - 1.0; e += p->kappa_v * d * d;
    d = s[2] / p->ref_curvature - 1.0; e += p->kappa_c * d * d;
    e += p->kappa_t * s[4];
    e += p->kappa_r * s[5];
    e += p->kappa_b * s[3];
    return e;
}

/* ------------ mesh_properties.cpp:51-148 + kernel.h:124-137 (gather gradient) */
void to_gradient(const to_mesh* m, const to_params* p, const to_nlist* nl, const double* vprops,
                 const double* sums, double* grad /*V*3*/) {
#pragma omp parallel for schedule(static)
    for (int idx = 0; idx < m->V; idx++) {
        vec3 dA = {{0, 0, 0}}, dV = dA, dC = dA, dB = dA, dT = dA;
        double c_to_a[3];
        c_to_a[0] = vprops[6 * idx + 2] / vprops[6 * idx + 0];
        int h0 = m->v_he[idx];
        if (h0 >= 0) {
            int he = h0;
            do {
                if (m->he_face[he] >= 0) {
                    int n_he = m->he_next[he];
                    int jdx = m->he_to[he], kdx = m->he_to[n_he];
                    c_to_a[1] = vprops[6 * jdx + 2] / vprops[6 * jdx + 0];
                    c_to_a[2] = vprops[6 * kdx + 2] / vprops[6 * kdx + 0];
                    double el = edge_length(m, he), ea = calc_dihedral_angle(m, he);
                    vec3 d_length = edge_length_grad0(m, he), d_angle = dihedral_angle_grad_bit0(m, he);
                    for (int i = 0; i < 2; i++) {
                        vec3 val = vscale(vadd(vscale(d_length, ea), vscale(d_angle, el)), 0.25);
                        dC = vadd(dC, val);
                        dB = vadd(dB, vscale(val, 4.0 * c_to_a[i]));
                    }
                    vec3 d_face_area = face_area_grad0(m, he);
                    for (int i = 0; i < 3; i++) {
                        vec3 val = vdiv(d_face_area, 3);
                        dA = vadd(dA, val);
                        dB = vsub(dB, vscale(val, 2.0 * c_to_a[i] * c_to_a[i]));
                    }
                    vec3 d_face_volume = face_volume_grad0(m, he);
                    for (int i = 0; i < 3; i++) dV = vadd(dV, vdiv(d_face_volume, 3));
                    el = edge_length(m, n_he);
                    d_angle = dihedral_angle_grad_bit2(m, n_he);
                    for (int i = 1; i < 3; i++) {
                        vec3 val = vscale(d_angle, 0.25 * el);
                        dC = vadd(dC, val);
                        dB = vadd(dB, vscale(val, 4.0 * c_to_a[i]));
                    }
                    dT = vadd(dT, bond_property_grad(m, p, he));
                }
                he = m->he_next[he ^ 1];
            } while (he != h0);
        }
        rep_ctx rc = {p, 0.0, {{0, 0, 0}}};
        if (nl) point_distances_cb(nl, m, idx, p->rep_lc1, rep_grad_cb, &rc);
        /* trimem_gradient (kernel.h:124-137) */
        double d, fac;
        vec3 g = {{0, 0, 0}};
        d = sums[0] / p->ref_area - 1.0; fac = 2.0 * p->kappa_a / p->ref_area * d; g = vadd(g, vscale(dA, fac));
        d = sums[1] / p->ref_volume - 1.0; fac = 2.0 * p->kappa_v / p->ref_volume * d; g = vadd(g, vscale(dV, fac));
        d = sums[2] / p->ref_curvature - 1.0; fac = 2.0 * p->kappa_c / p->ref_curvature * d; g = vadd(g, vscale(dC, fac));
        g = vadd(g, vscale(dT, p->kappa_t));
        g = vadd(g, vscale(rc.grad, p->kappa_r));
        g = vadd(g, vscale(dB, p->kappa_b));
        for (int c = 0; c < 3; c++) grad[3 * idx + c] = g.v[c];
    }
}

/* ------------------------------------ util.cpp:9-35 (area, edges_length) */
void to_area_edges(const to_mesh* m, int F, const int* f_he, double* area, double* length) {
    double a = 0.0, l = 0.0;
    for (int f = 0; f < F; f++) a += face_area(m, f_he[f]);
    for (int e = 0; e < m->HE / 2; e++) l += edge_length(m, 2 * e);
    *area = a; *length = l;
}

/* --------------------------- flips: OpenMesh is_flip_ok / flip, flip_utils.cpp:19-41 */
typedef struct {
    int V, HE, F;
    int *he_to, *he_next, *he_face, *v_he, *f_he;   /* mutable */
    const double* x;
} to_mmesh;

static to_mesh as_const(const to_mmesh* mm) {
    to_mesh m = {mm->V, mm->HE, mm->he_to, mm->he_next, mm->he_face, mm->v_he, mm->x};
    return m;
}

int to_is_flip_ok(const to_mmesh* mm, int e) {
    if (mm->he_face[2 * e] < 0 || mm->he_face[2 * e + 1] < 0) return 0;
    int ah = mm->he_to[mm->he_next[2 * e]], bh = mm->he_to[mm->he_next[2 * e + 1]];
    if (ah == bh) return 0;
    int h0 = mm->v_he[ah], h = h0;
    if (h0 >= 0) do { if (mm->he_to[h] == bh) return 0; h = mm->he_next[h ^ 1]; } while (h != h0);
    return 1;
}

void to_flip(to_mmesh* mm, int e) {
    int a0 = 2 * e, b0 = a0 + 1;
    int a1 = mm->he_next[a0], a2 = mm->he_next[a1], b1 = mm->he_next[b0], b2 = mm->he_next[b1];
    int va0 = mm->he_to[a0], va1 = mm->he_to[a1], vb0 = mm->he_to[b0], vb1 = mm->he_to[b1];
    int fa = mm->he_face[a0], fb = mm->he_face[b0];
    mm->he_to[a0] = va1; mm->he_to[b0] = vb1;
    mm->he_next[a0] = a2; mm->he_next[a2] = b1; mm->he_next[b1] = a0;
    mm->he_next[b0] = b2; mm->he_next[b2] = a1; mm->he_next[a1] = b0;
    mm->he_face[a1] = fb; mm->he_face[b1] = fa;
    mm->f_he[fa] = a0; mm->f_he[fb] = b0;
    if (mm->v_he[va0] == b0) mm->v_he[va0] = a1;
    if (mm->v_he[vb0] == a0) mm->v_he[vb0] = b1;
}

void to_edge_vertex_properties(const to_mmesh* mm, const to_params* p, const to_nlist* nl, int e, double* out) {
    to_mesh m = as_const(mm);
    double acc[6] = {0, 0, 0, 0, 0, 0}, t[6];
    for (int i = 0; i < 2; i++) {
        int heh = 2 * e + i;
        to_vertex_properties_one(&m, p, nl, m.he_to[heh], t);
        for (int q = 0; q < 6; q++) acc[q] += t[q];
        to_vertex_properties_one(&m, p, nl, m.he_to[m.he_next[heh]], t);
        for (int q = 0; q < 6; q++) acc[q] += t[q];
    }
    for (int q = 0; q < 6; q++) out[q] = acc[q];
}

/* flip_serial (flips.cpp:22-80) with injected proposals and uniform variates.
 * props: in = reduced properties of the mesh, out = after the sweep. Returns #accepted. */
int to_flip_sweep(to_mmesh* mm, const to_params* p, const to_nlist* nl, int n, const int* edges,
                  const double* u, double* props, int* accepted_flags, double* energies) {
    double e0 = to_energy(p, props);
    int acc = 0;
    for (int i = 0; i < n; i++) {
        int e = edges[i];
        if (accepted_flags) accepted_flags[i] = 0;
        if (to_is_flip_ok(mm, e)) {
            double oprops[6], nprops[6];
            to_edge_vertex_properties(mm, p, nl, e, oprops);
            for (int q = 0; q < 6; q++) props[q] -= oprops[q];
            to_flip(mm, e);
            to_edge_vertex_properties(mm, p, nl, e, nprops);
            for (int q = 0; q < 6; q++) props[q] += nprops[q];
            double en = to_energy(p, props), de = en - e0;
            double alpha = de < 0.0 ? 1.0 : exp(-de);
            if (u[i] <= alpha) {
                e0 = en; acc++;
                if (accepted_flags) accepted_flags[i] = 1;
            } else {
                to_flip(mm, e);
                for (int q = 0; q < 6; q++) { props[q] -= nprops[q]; props[q] += oprops[q]; }
            }
        }
        if (energies) energies[i] = e0;
    }
    return acc;
}
