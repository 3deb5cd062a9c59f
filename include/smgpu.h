/* smgpu.h -- C ABI of libsmgpu.so: the B200 (sm_100a) implementation of
 * smoothMesh's centroidal smoothing iteration.
 *
 * The reference (tkeskita/smoothMesh) has no FFI: its hot path is a set of
 * free functions called from main()'s loop, src/smoothMesh.C:2257-2437, on
 * OpenFOAM containers.  This header is the seam a maintainer binds instead
 * (see INTEGRATION.md for the OpenFOAM-side shim): plain pointers and sizes,
 * int status codes, no exceptions, no C++/torch types.
 *
 * Conventions
 *   - label = int32_t (OpenFOAM default), scalar = double, vector = 3 doubles
 *     (pointField storage), boolList = one uint8_t per element.
 *   - every function returns SMGPU_OK (0) or a negative error code;
 *     smgpu_last_error() returns the message of the last failure on the
 *     calling thread.  Conditions on which the reference calls
 *     FatalError/abort (src/smoothMesh.C:61-66, 354-362, 1073, 1087) are
 *     reported as SMGPU_ERR_MESH with the reference's message text.
 *   - one host thread per handle; calls on one handle must be serialised.
 *   - there is no CPU fallback: without a CUDA device smgpu_create fails.
 */
#ifndef SMGPU_H
#define SMGPU_H
#include <stdint.h>

#ifdef __cplusplus
extern "C"
{
#endif

#define SMGPU_OK 0
#define SMGPU_ERR_ARG -1
#define SMGPU_ERR_MESH -2
#define SMGPU_ERR_CUDA -3
#define SMGPU_ERR_COMM -4
#define SMGPU_ERR_IO -5

#define SMGPU_PATCH_BOUNDARY 0  /* wall, patch, ... : points on it are not "internal" */
#define SMGPU_PATCH_PROCESSOR 1 /* processorPolyPatch: skipped, src/smoothMesh.C:57-58 */
#define SMGPU_PATCH_EMPTY 2     /* emptyPolyPatch: fatal, src/smoothMesh.C:61-66 */

    /* What the host flattens an fvMesh into (replaces the `const fvMesh& mesh`
     * argument of every L3 function, e.g. src/smoothMesh.C:96-101). */
    typedef struct smgpu_mesh_desc
    {
        int64_t n_points, n_cells, n_faces, n_internal_faces;
        const double *points;        /* [3*n_points]  mesh.points()            */
        const int32_t *face_offsets; /* [n_faces+1]   mesh.faces() flattened   */
        const int32_t *face_verts;   /* [face_offsets[n_faces]]                */
        const int32_t *owner;        /* [n_faces]     mesh.faceOwner()         */
        const int32_t *neighbour;    /* [n_internal_faces] mesh.faceNeighbour() */
        int32_t n_patches;           /* mesh.boundaryMesh()                    */
        const int32_t *patch_start;  /* [n_patches]                            */
        const int32_t *patch_size;   /* [n_patches]                            */
        const int32_t *patch_kind;   /* [n_patches] SMGPU_PATCH_*              */
        /* decomposed (multi-GPU) runs only, else NULL: global label of every local
         * point (decomposePar's pointProcAddressing); defines which points are
         * shared between ranks for the syncTools::syncPointList replacements. */
        const int64_t *point_global_id;
        /* [n_patches] or NULL: 1 = patch selected by -layerPatches for the prismatic boundary
         * layer treatment (getPatchIdsForOption, src/smoothMesh.C:1823) */
        const int32_t *patch_layer;
    } smgpu_mesh_desc;

    /* The options of src/smoothMesh.C:1861-1918 that reach the hot path.
     * Negative min_edge_length / max_step_length select the reference defaults
     * (0.5 * mesh min edge length, 0.3 * min_edge_length; :1861-1865). */
    typedef struct smgpu_params
    {
        double min_edge_length;
        double max_step_length;
        double rel_step_frac;  /* 0.5  */
        double min_angle_deg;  /* 35   */
        double max_angle_deg;  /* 160  */
        double rel_tol;        /* 0.02 */
        int32_t total_min_freeze;      /* 0 */
        int32_t edge_angle_constraint; /* 1 */
        int32_t face_angle_constraint; /* 1 */
        int32_t geometry_variant;      /* 0 = openfoam.com face-centre formula, 1 = openfoam.org */
        int32_t device;                /* CUDA device ordinal */
        int32_t renumber;              /* 0 = keep the mesh's numbering; 1 = renumber points and cells along a Morton
                                          curve first (what OpenFOAM's renumberMesh would do before smoothMesh):
                                          gathers become local, results are those of the renumbered mesh
                                          (label-order-dependent tie-breaks / summation order follow the new
                                          labels), points and masks are returned in the caller's numbering */
        /* Boundary layer treatment (src/orthogonalBoundaryBlending.C; options :1892-1905), active when
         * some patch has patch_layer = 1 and layer_max_blending_fraction > 1e-15 (:2025).  On a processor mesh
         * the one-time set-up is collective and runs inside smgpu_comm_init (and smgpu_set_points). */
        double layer_max_blending_fraction; /* 0.3 */
        double layer_edge_length;           /* negative = min_edge_length (:1895-1896) */
        double layer_expansion_ratio;       /* 1.3 */
        int32_t min_layers;                 /* 1 */
        int32_t max_layers;                 /* 4 */
    } smgpu_params;

    typedef struct smgpu_handle smgpu_handle;

    void smgpu_default_params(smgpu_params *p);
    const char *smgpu_last_error(void);
    /* Boundary point smoothing (src/boundaryPointSmoothing.C): the arrays of
     * constant/geometry/initEdges.obj, targetEdges.obj (pass the initial edges again when that file is absent,
     * src/smoothMesh.C:2148-2160) and targetSurfaces.obj.  Edges are point pairs, triangles point triples. */
    typedef struct smgpu_boundary_geometry
    {
        int64_t n_init_points;
        const double *init_points;
        int64_t n_init_edges;
        const int32_t *init_edges;
        int64_t n_target_points;
        const double *target_points;
        int64_t n_target_edges;
        const int32_t *target_edges;
        int64_t n_surface_points;
        const double *surface_points;
        int64_t n_surface_tris;
        const int32_t *surface_tris;
        /* isCornerPoint / isFeatureEdgePoint label lists written by an earlier run (src/smoothMesh.C:2039-2078),
         * one entry per point, or NULL: when either holds a 1 the classes come from them instead of from the
         * initial edges, so a restart does not need init_edges to match the moved mesh */
        const int32_t *is_corner_point;
        const int32_t *is_feature_edge_point;
    } smgpu_boundary_geometry;
    /* Turns boundary point smoothing on for the patches with patch_smoothing[i] != 0 (-smoothingPatches): runs
     * the reference's one-time set-up (sanity checks :20-82, edge strings :557-590, classification :269-440,
     * hop counts and inner-neighbour map, point strings src/smoothMesh.C:2234-2250) on the mesh as it is now and
     * makes smgpu_iterate do :2307-2356 and the restore rule of :2387.  The target surface is searched through a
     * bounding volume hierarchy.  Call after smgpu_create (and again after smgpu_set_points).  On a processor mesh
     * the set-up is COLLECTIVE (global mesh figures, the two synchronised hop-count sweeps, the point normals): call
     * it on every rank after smgpu_comm_init and before smgpu_comm_p2p_connect (in-process groups:
     * smgpu_group_enable_boundary_smoothing); the per-iteration synchronisations of the feature (src/
     * boundaryPointSmoothing.C:660,668, orthogonalBoundaryBlending.C:491 for the inner map) ride in the interface
     * records of the iteration's one exchange.  Where the reference aborts the call returns SMGPU_ERR_MESH with its message. */
    int smgpu_enable_boundary_smoothing(smgpu_handle *h, const smgpu_boundary_geometry *geometry,
                                        const int32_t *patch_smoothing, double internal_smoothing_blending_fraction);
    /* counts of the boundary point classification after smgpu_enable_boundary_smoothing:
     * out = {corner points, feature edge points, smoothing surface points, target edge strings} */
    int smgpu_boundary_counts(smgpu_handle *h, int64_t out[4]);
    /* the classification as the two label lists the reference writes with the mesh (1 / 0 per point) */
    int smgpu_get_boundary_classes(smgpu_handle *h, int32_t *is_corner_point, int32_t *is_feature_edge_point);
    const char *smgpu_version(void);
    int smgpu_device_count(int32_t *n); /* visible CUDA devices (0 and SMGPU_ERR_CUDA if none) */

    /* Build derived connectivity, upload everything once, keep it resident in HBM. */
    int smgpu_create(const smgpu_mesh_desc *mesh, const smgpu_params *params, smgpu_handle **out);
    int smgpu_destroy(smgpu_handle *h);

    /* Effective parameters after default resolution, and mesh statistics
     * (getMeshStats, src/smoothMesh.C:1478-1541; findInternalMeshPoints :40-91). */
    int smgpu_get_params(smgpu_handle *h, smgpu_params *out);
    int smgpu_set_params(smgpu_handle *h, const smgpu_params *p);
    int smgpu_mesh_stats(smgpu_handle *h, double *min_edge, double *max_edge, int64_t *n_internal_points,
                         int64_t *n_edges);

    /* Run up to max_iters smoothing iterations (the loop body of
     * src/smoothMesh.C:2257-2437 minus file output).  Stops after the first
     * iteration whose residual < rel_tol (:2401).  n_frozen[i] / residual[i]
     * receive what the reference prints at :2396 for iteration i (either may be
     * NULL).  *iters_done = iterations executed.  In a multi-rank run the values
     * are the global sum / max on every rank. */
    int smgpu_iterate(smgpu_handle *h, int32_t max_iters, int64_t *n_frozen, double *residual, int32_t *iters_done);

    /* mesh.points() / isFrozenPoint of the last executed iteration. */
    int smgpu_get_points(smgpu_handle *h, double *points_out /* [3*n_points] */);
    int smgpu_set_points(smgpu_handle *h, const double *points_in /* [3*n_points] */);
    int smgpu_get_frozen(smgpu_handle *h, uint8_t *frozen_out /* [n_points] */);

    /* Device-timed duration (CUDA events) of the last smgpu_iterate call, in
     * milliseconds, and the number of kernels it launched. */
    int smgpu_last_timing(smgpu_handle *h, double *ms, int64_t *launches);

    /* Filter diagnostics of the last executed iteration: out = {1 if the face-angle filter runs
     * fused into the geometry tiles, points marked suspect by it (end points of edges it could not certify; they
     * take the literal evaluation), points active in the face-angle constraint (src/smoothMesh.C:1367-1368),
     * bit set: 1 geometry tiles, 2 second-generation tile kernel, 4 uniform hex fast path, 8 / 16 per-edge /
     * per-point single-precision filter level (global mirrors) in use, 32 per-point kernels on point tiles, 64 their
     * tile-local single-precision level}. */
    int smgpu_filter_stats(smgpu_handle *h, int64_t out[4]);

    /* Geometry tiles of this handle: out = {tiles (0 = two-kernel geometry), faces listed over all tiles (border
     * faces are listed by both tiles), points listed over all tiles, dynamic shared memory per block [bytes],
     * point tiles of the per-point kernels (0 = per-point kernels without tiles), points / cells listed over all
     * point tiles, dynamic shared memory of the edge-constraint tile kernel [bytes]}. */
    int smgpu_tile_stats(smgpu_handle *h, int64_t out[8]);

    /* Self-test of the shared-reciprocal division the geometry kernels use (one reciprocal for the three
     * components of a vector, the quotients bit-identical to IEEE division): about n random and structured
     * (a, d) triples on the device, *mismatches = components that differ from a / d (must be 0). */
    int smgpu_selftest_division(int32_t device, uint64_t seed, int64_t n, int64_t *mismatches);

    /* Optional per-kernel timing: when enabled, every kernel (group) launched by
     * smgpu_iterate is bracketed by CUDA events on the launch stream and the elapsed
     * times are accumulated per kernel name.  smgpu_profile() also clears the counters.
     * smgpu_profile_get: *n = number of kernel groups; arrays (any may be NULL) must hold
     * at least 16 entries; names are static strings. */
    int smgpu_profile(smgpu_handle *h, int32_t enable);
    int smgpu_profile_get(smgpu_handle *h, int32_t *n, const char **names, double *ms_total, int64_t *launches);

    /* ---- operator-level entry points -------------------------------------------
     * One call per reference L3 function, for parity tests and for hosts that
     * keep the reference's loop structure.  They operate on the handle's device
     * state: `points` (current mesh), `newPoints` (proposal), `isFrozenPoint`. */
    /* OpenFOAM geometry after movePoints: mesh.cellCentres() (src/smoothMesh.C:129,1218) */
    int smgpu_op_cell_centres(smgpu_handle *h, double *cell_centres_out /* [3*n_cells] or NULL */);
    /* isFrozenPoint = false (:2262), centroidalSmoothing (:96-166) + aspectRatioSmoothing
     * (:548-593) + constrainMaxStepLength (:684-754) -> newPoints */
    int smgpu_op_predict(smgpu_handle *h, double *new_points_out /* [3*n_points] or NULL */);
    /* restrictEdgeShortening (:602-652) then, if enabled, restrictMinEdgeAngleDecrease (:900-930) */
    int smgpu_op_edge_constraints(smgpu_handle *h, uint8_t *frozen_out /* or NULL */);
    /* restrictFaceAngleDeterioration (:1320-1437) */
    int smgpu_op_face_angle_constraint(smgpu_handle *h, uint8_t *frozen_out /* or NULL */);
    /* restore + count + calculateResidual + movePoints (:2384-2399) */
    int smgpu_op_commit(smgpu_handle *h, int64_t *n_frozen, double *residual);
    /* calculateBoundaryPointNormals (src/orthogonalBoundaryBlending.C:141-233), the call at :2266.  The reference
     * accumulates onto the previous call's normals (:178); this probe returns what the next call would produce and
     * leaves the handle's normals untouched.  The other smgpu_op_* entry points work on the handle's live state
     * (newPoints, isFrozenPoint, the face-angle work space) like the reference functions they stand for. */
    int smgpu_op_layer_normals(smgpu_handle *h, double *normals_out /* [3*n_points] or NULL */);
    /* updateNeighCoords + blendWithOrthogonalPoints + constrainMaxStepLength (:2283-2305) */
    int smgpu_op_layer_blend(smgpu_handle *h, double *new_points_out /* [3*n_points] or NULL */);
    /* calcMinMaxFaceAngleForEdge for the current mesh (:1135-1231): per-edge min/max */
    int smgpu_op_edge_face_angles(smgpu_handle *h, double *min_out /* [n_edges] */, double *max_out /* [n_edges] */);
    /* edge list (lo,hi) in the library's (OpenFOAM upper-triangular) numbering */
    int smgpu_get_edges(smgpu_handle *h, int32_t *edges_out /* [2*n_edges] */);
    /* derived tables, flattened: name in {pointCells, pointPoints, pointEdges, edgeFaces, edgeCells};
     * offsets may be NULL to query the number of values (returned through *n_values). */
    int smgpu_get_csr(smgpu_handle *h, const char *name, int32_t *offsets, int32_t *values, int64_t *n_values);

    /* ---- multi-GPU (one process per GPU) ------------------------------------------
     * Replaces syncTools::syncPointList / returnReduce over MPI (call sites
     * src/smoothMesh.C:134,142,402,429,455,472,2374,1567,2396) by NCCL over NVLink.
     * Every rank creates its handle from its processor mesh (decomposePar layout, with
     * point_global_id).  Start-up, in this order on every rank:
     *   1. smgpu_comm_local_shared: global labels of this rank's processor-patch points;
     *   2. the host all-gathers those lists (MPI / torch.distributed / anything) into
     *      counts[n_ranks] + the concatenation all_gids, and distributes the 128-byte
     *      ncclUniqueId made on rank 0 by smgpu_comm_unique_id;
     *   3. smgpu_comm_init builds the exchange plan and the NCCL communicator.
     * Afterwards smgpu_iterate runs the exchanges itself and returns global statistics. */
    int smgpu_comm_unique_id(uint8_t id_out[128]);
    int smgpu_comm_local_shared(smgpu_handle *h, int64_t *n, int64_t *gids_out /* or NULL */);
    int smgpu_comm_init(smgpu_handle *h, int32_t rank, int32_t n_ranks, const uint8_t nccl_unique_id[128],
                        const int64_t *counts, const int64_t *all_gids);

    /* Two-step form of the start-up for hosts that can agree on the outcome before anything collective runs:
     * smgpu_comm_prepare is purely local (exchange plan, buffers; fails on a mesh without point_global_id or on a
     * point shared by too many ranks), so the host can all-reduce its status and skip smgpu_comm_init on every
     * rank when any rank failed -- a rank that fails inside a one-step smgpu_comm_init would leave the others
     * waiting in ncclCommInitRank.  smgpu_comm_init after a successful prepare only does the collective half
     * (its counts / all_gids arguments are then ignored). */
    int smgpu_comm_prepare(smgpu_handle *h, int32_t rank, int32_t n_ranks, const int64_t *counts, const int64_t *all_gids);
    /* ---- peer-memory exchange (optional, after smgpu_comm_init) ----------------------------
     * Replaces the per-iteration NCCL calls (two grouped send/recv exchanges, two all-reduces) by stores into
     * peer memory issued from the kernels that produce the data, over NVLink: every rank owns one exchange block
     * (receive buffers, a flag word per neighbour and exchange, a statistics slot per rank); its peers map it --
     * CUDA IPC between processes, peer access between handles of one process -- and the producer kernels write
     * their records and then the flag with release semantics; the consumer kernels wait for the flags with
     * acquire loads.  Start-up: smgpu_comm_p2p_export on every rank, all-gather the 64-byte handles, then
     * smgpu_comm_p2p_connect with all of them (all_handles[64 * r] = rank r's; local_peers[r] non-NULL for ranks
     * that are handles of this process, which are then addressed directly; either argument may be NULL when the
     * other covers every rank).  All ranks must end up in the same mode: if connect fails anywhere, call
     * smgpu_comm_p2p_disable everywhere and the NCCL exchanges stay in use.  Results are identical in both modes. */
    int smgpu_comm_p2p_export(smgpu_handle *h, uint8_t handle_out[64]);
    int smgpu_comm_p2p_connect(smgpu_handle *h, const uint8_t *all_handles, smgpu_handle *const *local_peers);
    int smgpu_comm_p2p_disable(smgpu_handle *h);

    /* ncclCommAbort on this handle's communicator: releases kernels that wait for a peer which failed.  The
     * handle can only be destroyed afterwards. */
    int smgpu_comm_abort(smgpu_handle *h);

    /* ---- in-process group: several processor meshes on ONE device ----------------------
     * The same decomposed-case semantics (every member is one of the reference's MPI ranks) without NCCL: the
     * members are handles of this process on one device, driven by one host thread; the interface exchanges are
     * stream-ordered device copies between the members' buffers and the two reductions one small kernel.  Use:
     * a decomposed case with more processor directories than GPUs (the CLI's -parallel falls back to it), and
     * the parity tests of the exchange kernels on a single-GPU box.  handles[r] is rank r; every handle must be
     * freshly created from a processor mesh (with point_global_id) on the same device.  The group does not own
     * the handles: destroy the group first, then the handles.  Results per member through smgpu_get_points /
     * smgpu_get_frozen; smgpu_iterate on a member is refused. */
    typedef struct smgpu_group smgpu_group;
    int smgpu_group_create(smgpu_handle **handles, int32_t n, smgpu_group **out);
    int smgpu_group_iterate(smgpu_group *g, int32_t max_iters, int64_t *n_frozen, double *residual, int32_t *iters_done);
    int smgpu_group_destroy(smgpu_group *g);
    /* boundary point smoothing for all members (patch_smoothing[r] = member r's flags); see below */
    int smgpu_group_enable_boundary_smoothing(smgpu_group *g, const struct smgpu_boundary_geometry *geometry,
                                              const int32_t *const *patch_smoothing, double internal_smoothing_blending_fraction);

    /* Host-only helper (no GPU needed): the exchange plan smgpu_comm_init would build, as
     * flat arrays, so that host-side logic can be tested without devices.  local[i] / gids[i]
     * are this rank's processor-patch points.  Returns the number of send slots;
     * slot_point[slot] = local point, slot_rank[slot] = neighbour rank (both may be NULL). */
    int64_t smgpu_exchange_plan(int32_t rank, int32_t n_ranks, int64_t n_local, const int32_t *local,
                                const int64_t *gids, const int64_t *counts, const int64_t *all_gids,
                                int32_t *slot_point, int32_t *slot_rank);

#ifdef __cplusplus
}
#endif
#endif
