/* TEST INFRASTRUCTURE — plain-C restatement of LibGeoDecomp's per-timestep cell update for the
 * models in oracle/models/. It is a checker: only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load it. The product (libb200geo.so)
 * never links or calls it.
 *
 * Pinning: every function here is validated bit-for-bit against the reference's own
 * SerialSimulator built from /root/reference (oracle/_ref/lgd_ref_*, see oracle/Makefile)
 * by tests/test_oracle_pinning.py, and against the committed fixtures under tests/golden/
 * that those binaries generated (tests/golden/make_golden.py).
 *
 * All grids are dense, x fastest, [nz][ny][nx]; multi-member cells are member-major
 * (member m starts at byte offset cells * sum(sizeof(previous members))), which is the byte
 * stream SoAGrid::saveRegion produces for a box region (storage/soagrid.h:523-576,
 * lib/libflatarray/include/libflatarray/detail/save_functor.hpp:44-58).
 */
#ifndef B200GEO_ORACLE_H
#define B200GEO_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* kind: 6, 7 or 27 (oracle/models/jacobi.h). torus: 0 = Cube (edge constant), 1 = Torus. */
int oracle_jacobi(int kind, int torus, int nx, int ny, int nz, int steps, double edge,
                  const double *in, double *out);

/* Conway's Game of Life on 1-byte cells (oracle/models/conway.h). */
int oracle_gol(int torus, int nx, int ny, int steps, int edge_alive,
               const uint8_t *in, uint8_t *out);

/* D3Q19 BGK, float, 24 members (19 populations, density, velocityX/Y/Z, int state);
 * Cube topology, edge cell = LBMCellF() (oracle/models/lbm.h). Raw member-major in/out. */
int oracle_lbm(int nx, int ny, int nz, int steps, const void *in_raw, void *out_raw);

/* Member-major (de)serialisation of a list of streaks {x, y, z, endX}, restating
 * SoAGrid::saveRegion/loadRegion (storage/soagrid.h:523-576). grid_raw is a dense member-major
 * grid of nx*ny*nz cells; buf holds sum(streak lengths) cells, member-major. */
int oracle_save_region(int nx, int ny, int nz, int n_members, const int *member_bytes,
                       const void *grid_raw, const int *streaks, int n_streaks, void *buf);
int oracle_load_region(int nx, int ny, int nz, int n_members, const int *member_bytes,
                       void *grid_raw, const int *streaks, int n_streaks, const void *buf);

/* Short-range n-body in BoxCell<FixedArray<LJParticle<REAL>, cap>> containers on a Cube<3> grid of
 * nx*ny*nz cells of edge `edge` (oracle/models/nbody.h; storage/boxcell.h:112-174). counts:
 * int32 [nz][ny][nx]; parts: REAL [nz][ny][nx][cap][6] (pos xyz, vel xyz; unused slots zero on
 * output). real_bytes 4 or 8. Returns -3 when a container overflows (std::out_of_range in the
 * reference, storage/fixedarray.h:77-83). */
int oracle_nbody(int real_bytes, int nx, int ny, int nz, int cap, int steps, double dt, double cutoff, double edge,
                 const int32_t *counts_in, const void *parts_in, int32_t *counts_out, void *parts_out);
/* the same for a block of containers whose container (0,0,0) is container `origin` of a larger grid (container
 * (x, y, z) covers [(c + origin) * edge, + edge) per axis); everything outside the block is empty */
int oracle_nbody_at(int real_bytes, int nx, int ny, int nz, int cap, int steps, double dt, double cutoff, double edge,
                    const int origin[3], const int32_t *counts_in, const void *parts_in, int32_t *counts_out, void *parts_out);

/* ID-keyed mesh elements in ContainerCell<MeshElement, cap> containers (oracle/models/container.h;
 * storage/containercell.h:170-200, storage/neighborhoodadapter.h:45-65) on a Cube (torus = 0) or Torus grid of
 * nx*ny*nz containers, n_dims 2 (nz = 1) or 3. Arrays in the interchange format of include/b200geo.h (counts, ids,
 * values, influx, nb_counts, nb_ids [cells][cap][maxnb]); edge_*: the same six arrays for the ONE edge container, or
 * NULL for an empty one. values_out: double [cells][cap], zero for unused slots. Returns -2 and the ID in
 * *missing_id when an element lists an ID that none of the 3^n_dims containers around it holds (std::logic_error
 * "id not found" in the reference), -1 for bad arguments. */
int oracle_container(int n_dims, int torus, int nx, int ny, int nz, int cap, int maxnb, int steps,
                     const int32_t *counts, const int32_t *ids, const double *values, const double *influx,
                     const int32_t *nb_counts, const int32_t *nb_ids,
                     const int32_t *edge_count, const int32_t *edge_ids, const double *edge_values,
                     double *values_out, int32_t *missing_id);

#ifdef __cplusplus
}
#endif

#endif
