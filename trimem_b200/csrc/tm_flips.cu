// Edge flips on the device connectivity (placeholder until the flip kernels land).
#include "tm_common.cuh"

int tm_flips_release(tm_mesh* m) { (void)m; return 0; }

extern "C" {
int tm_flip_sweep(tm_mesh*, tm_nlist*, const tm_energy_params*, int, int, const int*, const double*,
                  int*, int*, double*) { tm_set_error("tm_flip_sweep: not implemented yet"); return 1; }
int tm_edge_vertex_properties(tm_mesh*, tm_nlist*, const tm_energy_params*, int, double*) {
    tm_set_error("tm_edge_vertex_properties: not implemented yet"); return 1; }
int tm_mesh_is_flip_ok(tm_mesh*, int, int*) { tm_set_error("tm_mesh_is_flip_ok: not implemented yet"); return 1; }
int tm_mesh_flip(tm_mesh*, int) { tm_set_error("tm_mesh_flip: not implemented yet"); return 1; }
}
