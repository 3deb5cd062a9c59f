/* smmesh.h -- C ABI of the host-side polyMesh helpers in libsmgpu.so:
 * reading/writing OpenFOAM polyMesh directories, synthetic mesh generators
 * (stand-ins for blockMesh / polyDualMesh, which are not available without
 * OpenFOAM) and cell decomposition (stand-in for decomposePar).
 *
 * Reference interfaces replaced: the fvMesh constructed by createMesh.H
 * (src/smoothMesh.C:1814-1818), mesh.write() (:2430), and the mesh
 * utilities the test scripts call (testcase/run_serial:11-16,
 * testcase/system/decomposeParDict).  Everything here is CPU-side setup; none
 * of it is on the per-iteration path.
 */
#ifndef SMMESH_H
#define SMMESH_H
#include <stdint.h>

#ifdef __cplusplus
extern "C"
{
#endif

    typedef struct smmesh smmesh;

    const char *smmesh_last_error(void);
    void smmesh_free(smmesh *m);

    /* constructors: return NULL on failure */
    smmesh *smmesh_gen_hex_block(int32_t nx, int32_t ny, int32_t nz, const double lo[3], const double hi[3]);
    /* brick `rank` = ix + px*(iy + py*iz) of a (nx*px) x (ny*py) x (nz*pz) block on [lo,hi], as a processor mesh */
    smmesh *smmesh_gen_hex_block_part(int32_t nx, int32_t ny, int32_t nz, int32_t px, int32_t py, int32_t pz,
                                      int32_t rank, const double lo[3], const double hi[3]);
    smmesh *smmesh_gen_kelvin(int32_t n, double h);
    /* brick `rank` (= ix + px*(iy + py*iz)) of that mesh as a processor mesh with point_global_id, generated
     * locally (BASELINE config 4 at its stated size is never materialised on one host) */
    smmesh *smmesh_gen_kelvin_part(int32_t n, double h, int32_t px, int32_t py, int32_t pz, int32_t rank);
    smmesh *smmesh_from_cells(int64_t n_points, const double *points, int32_t n_cells, const int32_t *cell_face_offsets,
                              const int32_t *cf_vert_offsets, const int32_t *cf_verts, const int32_t *cf_patch,
                              int32_t n_patches, const char *const *patch_names, const char *const *patch_types);
    smmesh *smmesh_from_arrays(int64_t n_points, const double *points, int64_t n_faces, const int32_t *face_offsets,
                               const int32_t *face_verts, const int32_t *owner, int64_t n_internal_faces,
                               const int32_t *neighbour, int64_t n_cells, int32_t n_patches, const int32_t *patch_start,
                               const int32_t *patch_size, const int32_t *patch_kind);
    smmesh *smmesh_read(const char *polymesh_dir);
    /* replace the point coordinates of m by those of an OpenFOAM points file (same point count) */
    int smmesh_read_points(smmesh *m, const char *points_file);
    int smmesh_write(const smmesh *m, const char *polymesh_dir, int32_t binary, int32_t precision);
    int smmesh_write_points(const double *points, int64_t n_points, const char *polymesh_dir, int32_t binary,
                            int32_t precision, const char *location);

    /* in-place: interior point jitter U(-amp, amp), counter-based RNG keyed on global point label */
    int smmesh_jitter(smmesh *m, double amp, uint64_t seed);

    /* sizes: what = 0 points, 1 cells, 2 faces, 3 internal faces, 4 face-vertex entries, 5 patches */
    int64_t smmesh_size(const smmesh *m, int32_t what);
    /* borrowed pointers, valid until smmesh_free */
    const double *smmesh_points(const smmesh *m);
    double *smmesh_points_mut(smmesh *m);
    const int32_t *smmesh_face_offsets(const smmesh *m);
    const int32_t *smmesh_face_verts(const smmesh *m);
    const int32_t *smmesh_owner(const smmesh *m);
    const int32_t *smmesh_neighbour(const smmesh *m);
    /* patch table: kind per smgpu.h SMGPU_PATCH_* */
    int smmesh_patches(const smmesh *m, int32_t *start, int32_t *size, int32_t *kind);
    const char *smmesh_patch_name(const smmesh *m, int32_t i);
    const int64_t *smmesh_point_global_id(const smmesh *m); /* NULL for undecomposed meshes */
    const int64_t *smmesh_cell_global_id(const smmesh *m);

    /* decomposed cases (decomposePar layout): processor<k>/constant/polyMesh/{points,faces,owner,neighbour,
     * boundary,pointProcAddressing,cellProcAddressing} */
    int smmesh_write_decomposed(smmesh *const *parts, int32_t n_parts, const char *case_dir, int32_t binary);
    smmesh *smmesh_read_processor(const char *case_dir, int32_t k);

    /* checkMesh-style quality of the mesh (what run_tests.sh:29,34 inspects after smoothing):
     * out = {max non-orthogonality [deg], average non-orthogonality [deg], max skewness, min angle between
     * consecutive face edges [deg], min edge length, max edge length, min cell volume}. */
    int smmesh_quality(const smmesh *m, double out[7]);

    /* Tiling of the fused geometry kernel (cells grouped along a space-filling curve, each tile with the list
     * of faces its cells touch; see smoothmesh_b200/csrc/topology.hpp GeomTiles), built and checked on the
     * host: out = {tiles (0 = a cell does not fit: two-kernel path), listed faces summed over tiles, largest
     * face list, faces of the mesh, largest point list, (edge, cell) pairs listed for the fused face-angle filter
     * (0 = some cell is not closed: per-edge kernel), edges per cell if uniform else 0, cells in uniform tiles (all
     * faces quadrilaterals, all cells hexahedra: the kernel's fast path)}.  Fails if an invariant the kernel relies
     * on does not hold. */
    int smmesh_geom_tiles(const smmesh *m, int32_t max_cells, int32_t max_faces, int32_t max_points, int64_t out[8]);

    /* One-time host set-up of boundary point smoothing (smoothmesh_b200/csrc/boundary.hpp; the reference's
     * classifyBoundaryPoints, findEdgeMeshStrings, calculatePointHopsToBoundary(smoothingPatches),
     * propagateInnerNeighInfo and the pointStrings loop, src/smoothMesh.C:2131-2250) for the mesh and the arrays
     * of constant/geometry/initEdges.obj / targetEdges.obj (edges = point pairs).  patch_smoothing: 0/1 per
     * patch.  Outputs (any may be NULL) are per point; corner_points is xyz.  Returns SMGPU_ERR_MESH with the
     * reference's FatalError text where the reference aborts. */
    int smmesh_boundary_setup(const smmesh *m, int64_t n_init_points, const double *init_points, int64_t n_init_edges,
                              const int32_t *init_edges, int64_t n_target_points, const double *target_points,
                              int64_t n_target_edges, const int32_t *target_edges, const int32_t *patch_smoothing,
                              double layer_edge_length, uint8_t *is_corner, uint8_t *is_feature_edge,
                              uint8_t *is_smoothing_surface, double *corner_points, int32_t *point_strings,
                              int32_t *hops_to_smoothing, int32_t *point_to_inner, int32_t *target_edge_strings);
    /* One-time host set-up of the boundary layer treatment for a serial mesh (topology.hpp: buildLayerSetup;
     * calculatePointHopsToBoundary and propagateOuterNeighInfo, src/orthogonalBoundaryBlending.C:52-134, :244-391):
     * hop counts (-1 = none), the point-to-outer-point map (-1 = none) and, per point, the boundary point whose
     * set-up normal it carries (-1 = zero normal).  patch_layer: 0/1 per patch. */
    int smmesh_layer_setup(const smmesh *m, const int32_t *patch_layer, int32_t max_layers, int32_t *hops,
                           int32_t *point_to_outer, int32_t *normal_source);

    /* Surface ray casts (the findLine of boundary point smoothing): nearest intersection of each segment
     * start[i] -> end[i] with the triangles, by visiting every triangle (use_bvh = 0, the reference definition
     * shared with the device kernel and the oracle) or through the bounding volume hierarchy of boundary.hpp
     * (use_bvh = 1), which must return the same triangle and the same point.  hit_tri[i] = -1 when there is none. */
    int smmesh_ray_cast(int64_t n_points, const double *points, int64_t n_tris, const int32_t *tris, int64_t n_rays,
                        const double *start, const double *end, int32_t use_bvh, int32_t *hit_tri, double *hit_point);

    /* labelIOList files next to the mesh (the isCornerPoint / isFeatureEdgePoint lists the reference keeps
     * between runs, src/smoothMesh.C:2039-2065): read returns the length (or -1 when the file is absent /
     * unreadable) and fills `data` when it is non-NULL; write emits OpenFOAM's `N{v}` form for uniform lists. */
    int64_t smmesh_read_label_list(const char *file, int32_t *data, int64_t capacity);
    int smmesh_write_label_list(const char *file, const char *object, const char *location, const int32_t *data,
                                int64_t n, int32_t binary);
    /* Wavefront OBJ reader used for constant/geometry: counts first (arrays NULL), then the data. */
    int smmesh_read_obj(const char *file, int64_t *n_points, double *points, int64_t *n_edges, int32_t *edges,
                        int64_t *n_tris, int32_t *tris);

    /* Morton (space-filling-curve) renumbering of points and cells, the renumberMesh stand-in: returns a new
     * valid polyMesh whose storage order keeps the smoothing kernels' gathers local.  The optional maps
     * receive the old label of every new point / cell.  Labels change, so label-order-dependent results are
     * those of the renumbered mesh (as if renumberMesh had been run before smoothMesh). */
    smmesh *smmesh_renumber(const smmesh *m, int32_t *point_old_of_new, int32_t *cell_old_of_new);

    /* decomposition: method 0 = bricks px*py*pz, 1 = recursive coordinate bisection into px parts.
     * parts_out receives n_parts (= px*py*pz or px) new meshes. */
    int smmesh_decompose(const smmesh *m, int32_t method, int32_t px, int32_t py, int32_t pz, smmesh **parts_out);

#ifdef __cplusplus
}
#endif
#endif
