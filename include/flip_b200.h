/*
 * flip_b200.h — C-ABI of the B200-native FLIP time step (libflip_b200.so).
 *
 * This is the drop-in boundary for the per-timestep hot path of the engine bundled in
 * jklae/FLIPEngine3D (src/engine, Blender-FLIP-Fluids 1.0.9).  Every entry point names the
 * reference interface it replaces (file:line relative to the reference tree).  Plain pointers and
 * sizes only; no C++/torch types.  All `float*`/`uint8_t*` arguments are HOST pointers unless the
 * name ends in `_dev`.  Every function returns FLIP_OK (0) or a nonzero status; the message is in
 * flip_last_error().  A context is single-caller, like FluidSimulation::update().
 *
 * Array layouts are the reference's (array3d.h:425-428, macvelocityfield.cpp:46-54):
 *   flat index i + W*(j + H*k);  U is (I+1,J,K), V (I,J+1,K), W (I,J,K+1);
 *   liquid phi (I,J,K) cell centred; solid phi (I+1,J+1,K+1) nodal; masks are 1 byte per entry.
 * Particles cross the boundary as the reference's MarkerParticle AoS {px,py,pz,vx,vy,vz}
 * (markerparticle.h:30-42) or as the two float-triplet blobs of loadMarkerParticleData
 * (fluidsimulation.h:102-106).
 */
#ifndef FLIP_B200_H
#define FLIP_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct flip_ctx flip_ctx;

enum {
    FLIP_OK = 0,
    FLIP_ERR_RUNTIME = 1,      /* std::runtime_error in the reference (e.g. update before initialize, fluidsimulation.cpp:5756) */
    FLIP_ERR_DOMAIN = 2,       /* std::domain_error (dt < 0, fluidsimulation.cpp:5761; bad dims/dx, :44-56) */
    FLIP_ERR_OUT_OF_RANGE = 3, /* std::out_of_range (bad index ranges, fluidsimulation.cpp:2059) */
    FLIP_ERR_CUDA = 4,         /* a CUDA call failed; there is NO CPU fallback */
    FLIP_ERR_UNSUPPORTED = 5   /* a feature outside the hot-path scope was requested */
};

/* Stage ids: the order of FluidSimulation::_stepFluid (fluidsimulation.cpp:5471-5508). */
enum {
    FLIP_STAGE_OBSTACLES = 0,     /* _updateObstacleObjects          :3196 (static scene: no-op)          */
    FLIP_STAGE_LIQUID_SDF = 1,    /* _updateLiquidLevelSet + postProcessSignedDistanceField :3225,:3246   */
    FLIP_STAGE_P2G = 2,           /* VelocityAdvector::advect         velocityadvector.cpp:38              */
    FLIP_STAGE_EXTRAPOLATE_A = 3, /* _extrapolateFluidVelocities      :3775 -> gridutils.cpp:32            */
    FLIP_STAGE_SAVE = 4,          /* _saveVelocityField               :3287                                */
    FLIP_STAGE_BODY_FORCE = 5,    /* _applyConstantBodyForces         :3450                                */
    FLIP_STAGE_PRESSURE = 6,      /* PressureSolver::solve            pressuresolver.cpp:44                */
    FLIP_STAGE_EXTRAPOLATE_B = 7, /* _extrapolateFluidVelocities      :3763                                */
    FLIP_STAGE_CONSTRAIN = 8,     /* _constrainVelocityFields         :3937                                */
    FLIP_STAGE_G2P = 9,           /* _updatePICFLIPMarkerParticleVelocities :4094                          */
    FLIP_STAGE_ADVANCE = 10,      /* _advanceMarkerParticles (+_removeMarkerParticles) :4355,:4324         */
    FLIP_STAGE_TAIL = 11,         /* _updateFluidObjects :4794 (nothing queued after load: no-op)          */
    FLIP_NUM_STAGES = 12
};

/* Array ids for flip_get_array / flip_set_array. */
enum {
    FLIP_ARRAY_U = 0, FLIP_ARRAY_V = 1, FLIP_ARRAY_W = 2,                 /* float */
    FLIP_ARRAY_VALID_U = 3, FLIP_ARRAY_VALID_V = 4, FLIP_ARRAY_VALID_W = 5, /* uint8 */
    FLIP_ARRAY_LIQUID_PHI = 6,  /* float (I,J,K)        ParticleLevelSet::_phi   particlelevelset.h:134 */
    FLIP_ARRAY_SOLID_PHI = 7,   /* float (I+1,J+1,K+1)  MeshLevelSet::_phi       meshlevelset.h:354     */
    FLIP_ARRAY_WEIGHT_U = 8, FLIP_ARRAY_WEIGHT_V = 9, FLIP_ARRAY_WEIGHT_W = 10, FLIP_ARRAY_WEIGHT_C = 11, /* WeightGrid pressuresolver.h:48 */
    FLIP_ARRAY_SAVED_U = 12, FLIP_ARRAY_SAVED_V = 13, FLIP_ARRAY_SAVED_W = 14,
    FLIP_ARRAY_NEAR_SOLID = 15, /* uint8 coarse grid, cell = 3dx  fluidsimulation.cpp:3094 */
    FLIP_ARRAY_PRESSURE = 16,   /* float (I,J,K) pressure of the last solve (pressuresolver.cpp:843-847) */
    FLIP_ARRAY_SOLID_VEL_U = 17, FLIP_ARRAY_SOLID_VEL_V = 18, FLIP_ARRAY_SOLID_VEL_W = 19, /* float, MAC layout: the solids'
                                   face velocities (VelocityDataGrid::field, meshlevelset.h:69-87); exist after
                                   flip_set_solid_velocity, and so does FLIP_ARRAY_WEIGHT_C */
    FLIP_NUM_ARRAYS = 20
};

/* Per-substep bookkeeping: the integers the reference logs in _logStepInfo (fluidsimulation.cpp:5710-5745). */
typedef struct flip_step_stats {
    int32_t particles;        /* "Fluid Particles"  after removal                                   */
    int32_t fluid_cells;      /* "Fluid Cells"      _getNumFluidCells :4761                        */
    int32_t pressure_rows;    /* PressureSolver::_matSize  pressuresolver.cpp:112                    */
    int32_t pcg_iterations;   /* "Pressure Solver Iterations"                                       */
    int32_t pcg_converged;    /* 1: reached tol; 2: hit max iterations but below acceptable tol; 0: failed */
    int32_t removed_solid;    /* particles deleted by _removeMarkerParticles, by reason             */
    int32_t removed_crowded;
    int32_t removed_fast;
    double pcg_error;         /* "Estimated Error"  = ||r||_inf                                      */
    double rhs_max;           /* ||b||_inf                                                           */
    double dt;                /* substep length                                                      */
} flip_step_stats;

/* ---- lifetime ---------------------------------------------------------------------------- */

/* FluidSimulation::FluidSimulation(int,int,int,double)  fluidsimulation.cpp:35-63.
 * `device` is the CUDA ordinal the context lives on. */
int flip_create(flip_ctx **out, int isize, int jsize, int ksize, double dx, int device);
/* FluidSimulation::~FluidSimulation  fluidsimulation.cpp:66 */
void flip_destroy(flip_ctx *ctx);
const char *flip_last_error(const flip_ctx *ctx);
/* status string of the last failed flip_create (ctx does not exist yet) */
const char *flip_create_error(void);

/* ---- configuration (before flip_initialize) ------------------------------------------------ */

/* FluidSimulation::addBodyForce(double,double,double)  fluidsimulation.cpp:1563 */
int flip_add_body_force(flip_ctx *ctx, double fx, double fy, double fz);
/* resetBodyForce :1598; enableExtremeVelocityRemoval / disableExtremeVelocityRemoval :1869-1881 (the speed rule of
 * _removeMarkerParticles :4345, on by default); setMarkerParticleScale :168-179 (the particle radius factor of the
 * surface reconstruction, 3.0 by default, :5083). */
int flip_reset_body_force(flip_ctx *ctx);
int flip_set_extreme_velocity_removal(flip_ctx *ctx, int on);
int flip_set_marker_particle_scale(flip_ctx *ctx, double scale);
/* FluidSimulation::setPICFLIPRatio :1440, setCFLConditionNumber :1100,
 * setMin/MaxTimeStepsPerFrame :1086-1098, pressure-solver members fluidsimulation.h:1657-1659 */
int flip_set_pic_flip_ratio(flip_ctx *ctx, double ratio);
int flip_set_cfl(flip_ctx *ctx, double cfl);
int flip_set_substep_limits(flip_ctx *ctx, int min_steps, int max_steps);
int flip_set_pressure_solver(flip_ctx *ctx, double tolerance, double acceptable_tolerance, int max_iterations);
/* ThreadUtils-free: the analogue of choosing the engine's parallel resources. 0 = Jacobi,
 * 1 = multigrid V-cycle (default). Both are GPU-parallel replacements of the reference's serial
 * MIC(0) (pcgsolver.h:69-221); the stopping rule is the reference's. */
int flip_set_preconditioner(flip_ctx *ctx, int kind);
/* Tuning of the multigrid V-cycle: damped-Jacobi sweeps before/after the coarse correction,
 * damping, weight of the coarse correction, sweeps on the coarsest level. */
int flip_set_multigrid(flip_ctx *ctx, int sweeps, double damping, double coarse_weight, int coarsest_sweeps);
/* Per-sweep damping of the `sweeps` pre-smoothing passes (the post-smoothing passes use the mirrored order,
 * which keeps the V-cycle symmetric as CG requires).  Reciprocals of the Chebyshev nodes of the interval of
 * the Jacobi-scaled spectrum to be damped turn the sweeps into a polynomial smoother at no extra cost. */
int flip_set_multigrid_schedule(flip_ctx *ctx, int sweeps, const double *damping);
/* Initial guess of the PCG.  0: zero, as the reference (pcgsolver.h:258).  1 (default): the pressure of the
 * previous substep on the cells that were pressure rows then (zero elsewhere).  The stopping rule is unchanged
 * (||r||_inf <= tol * ||b||_inf with r = b - A x), so the answer agrees to the solver tolerance; only the
 * number of iterations drops. */
int flip_set_pressure_warm_start(flip_ctx *ctx, int on);
/* 1: the PCG solve runs as one persistent cooperative kernel (device-side iteration loop,
 * grid-wide barriers); 0 (default): one launch per solver pass, host polls the convergence flag. Same arithmetic. */
int flip_set_solver_mode(flip_ctx *ctx, int persistent);

/* Arithmetic of the trilinear MAC sampling in G2P and RK3 (MACVelocityField::evaluateVelocityAtPositionLinear,
 * macvelocityfield.cpp:519-645).  FLIP_SAMPLING_EXACT restates the reference's double-precision blend
 * operation by operation: particle velocities and positions come out bit-identical to the reference's from
 * identical inputs.  FLIP_SAMPLING_FAST (default) keeps the cell indices and interpolation weights exact
 * (they are exact in float when dx is a power of two) and evaluates only the 8-point blend in float: results
 * agree to a few ulp (rel-L2 ~1e-7, inside the 1e-4 contract) at a fraction of the instruction count.  Falls
 * back to EXACT per sample outside the interior of the grid and entirely when dx is not a power of two.
 * The same switch selects the P2G splat weights (VelocityAdvector, velocityadvector.cpp:437): EXACT evaluates
 * the weight polynomial literally (left to right, scalar), so face values differ from the reference by float
 * summation order only; FAST evaluates it in Horner form with packed FP32 pairs (weights agree to < 1e-6, face
 * values to rel-L2 ~1e-7; faces whose total weight is near the 1e-6 validity threshold are recomputed
 * literally, so the valid masks stay exact).  The liquid SDF is bit-exact in both modes. */
enum { FLIP_SAMPLING_EXACT = 0, FLIP_SAMPLING_FAST = 1 };
int flip_set_sampling_mode(flip_ctx *ctx, int mode);

/* FluidSimulation::loadMarkerParticleData  fluidsimulation.cpp:2488 — float xyz triplets; copied;
 * applied (with the in-domain filter of _loadMarkerParticles :2773) at flip_initialize(). */
int flip_load_particles(flip_ctx *ctx, int n, const float *positions_xyz, const float *velocities_xyz);
/* FluidSimulation::addMeshFluid(MeshObject, velocity)  fluidsimulation.cpp:1573-1590: queues a fluid object.  Like the
 * reference's queue (_addedFluidMeshObjectQueue) it is seeded at the END of the next substep (_updateFluidObjects :5504
 * -> _updateAddedFluidMeshObjectQueue :4724-4759) -- on the device (csrc/seed.cu): eight points (+-dx/4)^3 per cell that
 * has a corner node inside the object (MeshObject::getCells, meshobject.cpp:101-140), kept where the object's signed
 * distance (trilinear sample of its nodal field) is <= 0, the solid SDF is > 0 and the sub-cell holds no particle yet
 * (ParticleMaskGrid, particlemaskgrid.cpp:55-112).  Until it is seeded its |velocity| enters the predicted maximum
 * speed of the first time step (_predictMaximumMarkerParticleSpeed :5529-5551).  Not reproduced: the reference's jitter
 * of 2.5e-4 dx from the unseeded rand().
 *   flip_add_fluid_sdf: the object as a NODAL signed distance field of the global grid, (I+1)(J+1)(K+1) floats, negative
 *     inside -- what MeshLevelSet::fastCalculateSignedDistanceField leaves in _phi (copied); cell_lo / cell_hi: the cell
 *     range [lo,hi) to scan (NULL: the whole grid).
 *   flip_add_fluid_box: an axis-aligned box [lo,hi] of world coordinates (the object FluidManager builds,
 *     src/FluidManager.cpp:56-64); its nodal distances are evaluated in place. */
int flip_add_fluid_sdf(flip_ctx *ctx, const float *nodal_sdf, const int cell_lo[3], const int cell_hi[3], const double velocity[3]);
int flip_add_fluid_box(flip_ctx *ctx, const double lo[3], const double hi[3], const double velocity[3]);
/* FluidSimulation::addMeshFluidSource / removeMeshFluidSource  fluidsimulation.cpp:1953-1985 with a MeshFluidSource
 * (meshfluidsource.h) that is an inflow (setInflow: emits at the end of EVERY substep where sub-cells are free,
 * _updateInflowMeshFluidSource :4561-4604 with the default substep emissions of 1) or an outflow (setOutflow + fluid
 * outflow: removes the particles inside it, _updateOutflowMeshFluidSource :4607-4668, not inversed).  Static sources,
 * given as a box or as a nodal signed distance field like the queued objects above; *id identifies the source for
 * flip_enable_fluid_source (MeshFluidSource::enable / disable) and flip_remove_fluid_source.  Same device kernels as
 * the queue (csrc/seed.cu). */
int flip_add_fluid_source_box(flip_ctx *ctx, int outflow, const double lo[3], const double hi[3], const double velocity[3], int *id);
int flip_add_fluid_source_sdf(flip_ctx *ctx, int outflow, const float *nodal_sdf, const int cell_lo[3], const int cell_hi[3],
                              const double mesh_lo[3], const double mesh_hi[3], const double velocity[3], int *id);
/* mesh_lo / mesh_hi: the bounding box of the source's mesh (may be NULL: the box of the nodes with sdf <= 0).  It fixes
 * the extent of the level-set grid the reference keeps per source (MeshFluidSource::update, meshfluidsource.cpp:178-198),
 * which the inflow velocity constraint reads at un-offset world positions (fluidsimulation.cpp:3401, :4138) -- a quirk
 * of the reference that decides where the constraint acts and that is reproduced as it is. */
int flip_enable_fluid_source(flip_ctx *ctx, int id, int on);
int flip_remove_fluid_source(flip_ctx *ctx, int id);
/* MeshFluidSource::enableConstrainedFluidVelocity / disableConstrainedFluidVelocity (meshfluidsource.cpp:146-152; on by
 * default): while on, the faces inside an inflow get no body force (_getInflowConstrainedVelocityComponents
 * fluidsimulation.cpp:3372) and the particles inside it carry the source's velocity after the PIC/FLIP update
 * (_constrainMarkerParticleVelocities :4113). */
int flip_constrain_fluid_source_velocity(flip_ctx *ctx, int id, int on);
/* Static obstacles: FluidSimulation::addMeshObstacle / removeMeshObstacle (fluidsimulation.cpp:1994-2031) and
 * MeshObject::enable / disable for meshes that do not move.  The obstacle comes as an axis-aligned box or as the nodal
 * signed distance field a MeshLevelSet holds for it ((I+1)(J+1)(K+1) floats, negative inside, a large positive value
 * where the field was not computed); it is merged into the solid SDF by minimum, as MeshLevelSet::calculateUnion does
 * (meshlevelset.cpp:1758-1795), and the face weights, the near-solid mask and everything that reads the solid SDF
 * (collision, removal, seeding, the surface clamp) follow.  Before flip_initialize the merge happens there; afterwards at
 * the start of the next substep (the reference's _isSolidLevelSetUpToDate = false, :2007).  Obstacles have zero velocity
 * and zero friction (MeshObject defaults) unless flip_set_solid_velocity supplies face velocities; the per-substep SDF of an
 * animated / rigid-body obstacle is not built by the library.
 * flip_remove_obstacle of an unknown id: FLIP_ERR_DOMAIN (the reference throws std::invalid_argument). */
int flip_add_obstacle_box(flip_ctx *ctx, const double lo[3], const double hi[3], int *id);
int flip_add_obstacle_sdf(flip_ctx *ctx, const float *nodal_sdf, int *id);
int flip_enable_obstacle(flip_ctx *ctx, int id, int on);
int flip_remove_obstacle(flip_ctx *ctx, int id);
/* Moving solids, the hot-path half: the face velocities of the solids, as MeshLevelSet keeps them
 * beside the solid SDF (getFaceVelocityU/V/W, meshlevelset.cpp:207-231; the reference fills them from the vertex
 * velocities of animated meshes and normalises them, :640-702, :1319-1372).  Three HOST arrays in the MAC layout
 * (U (I+1)JK, V I(J+1)K, W IJ(K+1)), copied; three NULLs: every solid at rest again (the default).  While set,
 *   - the right-hand side of the pressure system carries the solid terms +-(w_face - w_centre) u_solid / dx of
 *     _calculateNegativeDivergenceVectorThread (pressuresolver.cpp:595-613), w_centre being the cell-centre entry of the
 *     weight grid (_updateWeightGridThread CENTER :3720-3727 over MeshLevelSet::_getCellWeight, meshlevelset.cpp:1490-1513),
 *     which is built from the solid SDF from then on (FLIP_ARRAY_WEIGHT_C);
 *   - before every solve the velocities around liquid regions that are enclosed by solids and touch no air are set to
 *     zero IN the stored arrays, as _conditionSolidVelocityField does (pressuresolver.cpp:124-244; regions of one cell
 *     are left alone, :219), on the device;
 *   - faces of zero weight take the solid's face velocity in the solid constraint (_constrainVelocityFieldThread
 *     fluidsimulation.cpp:3884-3933; partly open faces follow the friction, see flip_set_boundary_friction).
 * The SDF of a moving solid itself is the caller's (flip_add_obstacle_sdf / flip_enable_obstacle / flip_remove_obstacle per
 * substep, or flip_set_solid_sdf before flip_initialize).  Not available in a z-slab run (FLIP_ERR_UNSUPPORTED). */
int flip_set_solid_velocity(flip_ctx *ctx, const float *U, const float *V, const float *W);
/* Animated obstacles, rigid translation of a box: MeshObject::updateMeshAnimated(previous, current, next)
 * (meshobject.cpp:61-95) for an obstacle added with flip_add_obstacle_box -- the three meshes are that box moved by the
 * three offsets; call it once per frame before flip_update, as the reference's callers do.  While such an obstacle is
 * enabled, the obstacle stage of EVERY substep (_updateSolidLevelSet fluidsimulation.cpp:3028-3071 rebuilds the solid SDF
 * while an animated mesh changes, :3002) places the box at current + t (next - current) with t the part of the frame
 * completed at the start of the substep (:2892-2893; MeshObject::getMesh(t) meshobject.cpp:158-177), gives it the
 * velocity ((current - previous) + t ((next - current) - (current - previous))) / frame dt (getVertexVelocities
 * :199-215), re-derives solid SDF, face / centre weights and near-solid mask, and builds the solids' face velocities the
 * way the merged MeshLevelSet does: per face the sum over solids of solid fraction x velocity divided by the summed
 * solid fractions where that exceeds 1e-6 (_computeVelocityGridThread meshlevelset.cpp:1319-1372, calculateUnion
 * :1797-1828, _normalizeVelocityGridThread :1738-1756), then 5 layers of extrapolation (:697, the fluid's routine, on
 * the device).  The per-substep SDF and the solid fractions are HOST work (as in the reference); divergence,
 * conditioning, constraint and extrapolation run on the device (flip_set_solid_velocity describes them; the arrays it
 * would set are overwritten every substep while an animated obstacle is enabled).  General animated meshes:
 * flip_add_obstacle_mesh / flip_set_obstacle_mesh_motion below.  Not available in a z-slab run. */
int flip_set_obstacle_box_motion(flip_ctx *ctx, int id, const double offset_prev[3], const double offset_cur[3],
                                 const double offset_next[3]);
/* Animated obstacles, general closed meshes of fixed topology (rigid or deforming): flip_add_obstacle_mesh hands the library
 * the mesh itself (addMeshObstacle with a MeshObject; static until animated: its signed distance field is flip_mesh_sdf's),
 * flip_set_obstacle_mesh_motion the vertices of the previous, the current and the next frame (updateMeshAnimated; same
 * count and order as the mesh was added with), once per frame.  Per substep as above, with per-vertex velocities; the
 * face velocity is that of the mesh surface nearest to the face centre, found and interpolated as the reference does
 * (MeshLevelSet::getNearestVelocity meshlevelset.cpp:168-205 over the closest triangles of the surrounding nodes,
 * _pointToTriangleVelocity :1567-1640).  flip_mesh_velocity_data is that computation on its own (HOST code, no device
 * needed; the CPU tests pin it to the reference): phi as flip_mesh_sdf, per face of U / V / W the mesh's solid fraction
 * and fraction x velocity component (any output but phi may be NULL). */
int flip_add_obstacle_mesh(flip_ctx *ctx, const float *vertices_xyz, int num_vertices, const int *triangles, int num_triangles, int *id);
int flip_set_obstacle_mesh_motion(flip_ctx *ctx, int id, const float *vertices_prev, const float *vertices_cur, const float *vertices_next);
int flip_mesh_velocity_data(int isize, int jsize, int ksize, double dx, const float *vertices_xyz, int num_vertices, const int *triangles,
                            int num_triangles, const float *vertex_velocities_xyz, int band, float far_value, float *phi,
                            float *fractionU, float *fractionV, float *fractionW, float *fieldU, float *fieldV, float *fieldW);
/* Friction of the solids: FluidSimulation::setBoundaryFriction (fluidsimulation.cpp:1747-1759; std::domain_error outside
 * [0,1]) and MeshObject::setFriction of an obstacle (meshobject.cpp:300-304; clamped to [0,1]); both 0 by default.  In the
 * solid constraint a partly open face (0 < weight < 1) becomes f u_solid + (1 - f) u (_constrainVelocityFieldThread
 * :3895-3900) with the face friction f of _getFaceFrictionU/V/W (:3785-3853): a quarter of the sum over the face's four
 * nodes of the friction of the solid the merged level set names as closest there (MeshLevelSet::getClosestMeshObject; the
 * merge order and take-over rule of calculateUnion, meshlevelset.cpp:1782-1796, are restated on the host).  Re-derived
 * with the static inputs (at flip_initialize, after a change at the next substep, every substep with animated obstacles).
 * flip_set_face_friction hands the three face arrays in directly (MAC layout, copied; NULLs: derived again) and
 * flip_get_face_friction reads back what the constraint uses (zeros while every friction is 0).  Not in z-slab runs. */
int flip_set_boundary_friction(flip_ctx *ctx, double friction);
int flip_set_obstacle_friction(flip_ctx *ctx, int id, double friction);
int flip_set_face_friction(flip_ctx *ctx, const float *U, const float *V, const float *W);
int flip_get_face_friction(flip_ctx *ctx, float *U, float *V, float *W);
/* HOST utilities behind the above (no CUDA device needed; the CPU tests pin them to the reference): the face friction of
 * num_solids solids given as nodal fields in merge order -- phis[0] the domain (defined everywhere), the others obstacle
 * fields carrying the largest float outside their band, as flip_box_obstacle_sdf / flip_mesh_sdf(far = FLT_MAX) produce
 * them -- with their frictions; and the nodal field of a box obstacle. */
int flip_face_friction(int isize, int jsize, int ksize, double dx, int band, int num_solids, const float *const *phis,
                       const float *frictions, float *fU, float *fV, float *fW);
int flip_box_obstacle_sdf(int isize, int jsize, int ksize, double dx, int band, const double lo[3], const double hi[3], float *phi);
/* The cell-centre weights of a nodal solid SDF, computed on the HOST exactly as the library computes them (no CUDA device
 * needed): phi (I+1)(J+1)(K+1) floats in, wC IJK floats out. */
int flip_center_weights(int isize, int jsize, int ksize, double dx, const float *phi_nodal, float *wC);
/* FluidSimulation::_addMarkerParticle  fluidsimulation.cpp:2637 (range-checked push). */
int flip_add_marker_particle(flip_ctx *ctx, const float position[3], const float velocity[3]);

/* The static inputs of a box domain, computed on the HOST exactly as flip_initialize computes them (no CUDA device
 * needed; host code of csrc/static_host.cpp): the nodal solid SDF of the reference's domain box
 * (_addStaticObjectsToSDF fluidsimulation.cpp:2927-2936, box of :2834-2839), the face weights
 * (_updateWeightGridThread :3690-3730) and the coarse near-solid mask (:3083-3125).  Any output pointer may be NULL.
 * phi: (I+1)(J+1)(K+1) floats; wU (I+1)JK, wV I(J+1)K, wW IJ(K+1) floats; near_solid: ceil(I/3)*ceil(J/3)*ceil(K/3)
 * bytes, its dimensions returned in near_dims[3].  phi_is_input != 0: phi is READ (as flip_set_solid_sdf would supply it)
 * and the weights / mask are derived from it instead of from the built-in box. */
int flip_static_inputs(int isize, int jsize, int ksize, double dx, float *phi, int phi_is_input, float *wU, float *wV, float *wW,
                       unsigned char *near_solid, int near_dims[3]);

/* Override the static solid inputs.  By default the context builds the reference's
 * domain box (inset 1.5dx+5e-5, fluidsimulation.cpp:2834-2839) itself.  phi: (I+1)(J+1)(K+1) floats. */
int flip_set_solid_sdf(flip_ctx *ctx, const float *phi_nodal);
/* The nodal signed distance field of a closed triangle mesh on the grid, as a MeshLevelSet holds it after
 * fastCalculateSignedDistanceField(mesh, band) (meshlevelset.cpp:773-828; MeshObject::getMeshLevelSet meshobject.cpp:239):
 * exact distances at the nodes within `band` cells of the mesh's index box, negative inside, far_value (or (band+1) dx when
 * far_value <= 0) elsewhere.  HOST code, no device needed; the input of flip_add_fluid_sdf / flip_add_fluid_source_sdf /
 * flip_add_obstacle_sdf for meshes that are not axis-aligned boxes.  phi: (I+1)(J+1)(K+1) floats; cell_lo / cell_hi
 * (may be NULL): the cell range worth scanning. */
int flip_mesh_sdf(int isize, int jsize, int ksize, double dx, const float *vertices_xyz, int num_vertices, const int *triangles,
                  int num_triangles, int band, float far_value, float *phi, int cell_lo[3], int cell_hi[3]);

/* FluidSimulation::initialize  fluidsimulation.cpp:82 */
int flip_initialize(flip_ctx *ctx);

/* ---- the hot path -------------------------------------------------------------------------- */

/* FluidSimulation::update(double dt)  fluidsimulation.cpp:5755 — one frame, CFL substeps inside. */
int flip_update(flip_ctx *ctx, double dt);
/* FluidSimulation::getCurrentFrame :94 */
int flip_get_current_frame(const flip_ctx *ctx, int *frame);
/* FluidSimulation::setCurrentFrame  fluidsimulation.cpp:98 — with flip_get_particle_positions / _velocities and
 * flip_load_particles this is the reference's snapshot / restore contract: a context
 * restored from the two float-triplet blobs and the frame number continues the run (frame 0 alone uses the
 * predicted first time step, :5595). */
int flip_set_current_frame(flip_ctx *ctx, int frame);
/* number of substeps the last flip_update took, and their stats (index 0..substeps-1) */
int flip_get_num_substeps(const flip_ctx *ctx, int *substeps);
int flip_get_step_stats(const flip_ctx *ctx, int substep, flip_step_stats *out);

/* FluidSimulation::getNumMarkerParticles :2049 */
int flip_get_num_particles(const flip_ctx *ctx, int *n);
/* FluidSimulation::getMarkerParticles() :2053 — copies n*6 floats {p,v} to host. */
int flip_get_particles(flip_ctx *ctx, float *aos6, int capacity);
/* Replace the particle store in place (the reverse of flip_get_particles; lock-step tests and the
 * end-to-end benchmark leg use it).  Applies the in-domain filter of _addMarkerParticle. */
int flip_set_particles(flip_ctx *ctx, int n, const float *aos6);
/* FluidSimulation::getMarkerParticlePositionData / VelocityData :2408-2420 (xyz triplets). */
int flip_get_particle_positions(flip_ctx *ctx, float *xyz, int capacity);
int flip_get_particle_velocities(flip_ctx *ctx, float *xyz, int capacity);

/* Optional particle identity: when enabled (before particles are uploaded), every particle carries the
 * index it had in the last flip_load_particles / flip_set_particles call through the per-step cell
 * sort, so callers can match particles across steps although the store is reordered (the reference
 * keeps insertion order, fragmentedvector.h; here order is by cell).  Costs 8 B/particle/step. */
int flip_enable_particle_ids(flip_ctx *ctx, int on);
int flip_get_particle_ids(flip_ctx *ctx, int32_t *ids, int capacity);
/* ids count from `base` instead of 0 (before flip_initialize): a z-slab rank that loads only its own part of a scene
 * passes the number of particles that precede it, so that ids are global. */
int flip_set_particle_id_base(flip_ctx *ctx, int base);

/* FluidSimulation::getIsomesh()  fluidsimulation.h:1126 -- the surface FluidManager draws (src/FluidManager.cpp:102-171):
 * ParticleMesher::meshParticles (particlemesher.cpp:36: scalar field of the particles on the grid subdivided
 * setSurfaceSubdivisionLevel times :228-231, marching cubes polygonizer3d.cpp:643-676) and TriangleMesh::smooth
 * (trianglemesh.cpp:536-583; value 0.5, 2 iterations by default, fluidsimulation.h:1607-1608), at the engine's other
 * defaults; reconstructed on the device (csrc/mesher.cu) from the particles as they stand at the call and cached until
 * the next step.  flip_get_isomesh_size first, then flip_get_isomesh with room for 3 floats per vertex and 3 vertex
 * indices per triangle.  Single GPU (a z-slab run gathers its particles on one rank first). */
int flip_set_surface_subdivision_level(flip_ctx *ctx, int n);
int flip_set_surface_smoothing(flip_ctx *ctx, double value, int iterations);
int flip_get_isomesh_size(flip_ctx *ctx, int *num_vertices, int *num_triangles);
int flip_get_isomesh(flip_ctx *ctx, float *vertices_xyz, int *triangles);
/* parity seam: the mesher's scalar field on the (I s+1)(J s+1)(K s+1) nodes of the subdivided grid, i fastest: the
 * inside flag of every node (value > 0 and not solid), whether the node belongs to a surface cell (`need`), and the
 * exact value of those nodes (0 elsewhere).  Any pointer may be NULL. */
int flip_get_isomesh_field(flip_ctx *ctx, float *values, unsigned char *inside, unsigned char *need);
/* the marching-cubes case table the reconstruction uses (generated on the host, csrc/mc_tables.h; no CUDA device
 * needed): per sign configuration (bit c = corner c inside; corner bits x | y << 1 | z << 2) the triangle count and up
 * to 8 triples of crossed edges (edge = axis * 4 + the bits of its lower corner on the two other axes). */
int flip_mc_case_table(unsigned char counts[256], unsigned char edge_triples[256 * 24]);

/* FluidSimulation::getVelocityField :2242 + MACVelocityField::getRawArrayU/V/W macvelocityfield.cpp:100-110 */
int flip_get_velocity_field(flip_ctx *ctx, float *U, float *V, float *W);

/* ---- per-stage seams (the reference's *Parameters structs: velocityadvector.h:62-67,
 *      pressuresolver.h:63-79) for isolated parity tests and profiling ------------------------ */

/* Frame/substep bookkeeping of update() (fluidsimulation.cpp:5768-5824) exposed step by step. */
int flip_begin_frame(flip_ctx *ctx, double dt);
int flip_begin_substep(flip_ctx *ctx, double *dt_substep);
int flip_run_stage(flip_ctx *ctx, int stage, double dt_substep);
int flip_end_substep(flip_ctx *ctx, int *more);
int flip_end_frame(flip_ctx *ctx);

int flip_array_bytes(const flip_ctx *ctx, int which, int64_t *bytes);
int flip_get_array(flip_ctx *ctx, int which, void *host_out);
int flip_set_array(flip_ctx *ctx, int which, const void *host_in);

/* Device time of each stage in the last substep, milliseconds (CUDA events on the context's stream) —
 * the analogue of the reference's TimingData buckets (fluidsimulation.h:1172-1233). */
int flip_get_stage_times_ms(const flip_ctx *ctx, float ms[FLIP_NUM_STAGES]);
/* Per-kernel-class device time, measured with CUDA event pairs recorded on the context's stream
 * around the launches of each class while enabled.  total_ms / launches = mean launch duration. */
enum {
    FLIP_KERNEL_SDF_P2G = 0,   /* fused liquid-SDF + P2G gather                         */
    FLIP_KERNEL_G2P = 1,       /* PIC/FLIP grid-to-particle                             */
    FLIP_KERNEL_ADVANCE = 2,   /* RK3 + collision                                       */
    FLIP_KERNEL_SORT = 3,      /* whole cell sort + removal pipeline (several launches) */
    FLIP_KERNEL_EXTRAPOLATE = 4, /* one extrapolateVelocityField call (3 components)    */
    FLIP_KERNEL_PCG_SPMV = 5,  /* one operator application (+ fused dot)                */
    FLIP_KERNEL_PCG_ITER = 6,  /* one whole PCG iteration                               */
    FLIP_KERNEL_PRESSURE_BUILD = 7, /* row enumeration + rhs + matrix                   */
    FLIP_KERNEL_PRESSURE_APPLY = 8, /* velocity update                                  */
    FLIP_KERNEL_PRECOND = 9,   /* one preconditioner application (multigrid V-cycle), multi-launch solver only */
    FLIP_KERNEL_PCG_SOLVE = 10, /* the whole PCG solve as one persistent cooperative kernel */
    FLIP_KERNEL_PCG_DIR_SPMV = 11, /* search-direction update of iteration it-1 fused with the operator application of
                                      iteration it (single-GPU multigrid PCG)                */
    FLIP_KERNEL_G2P_ADVANCE = 12,  /* PIC/FLIP update fused with RK3 + collision              */
    FLIP_NUM_KERNEL_CLASSES = 13
};
int flip_enable_kernel_timing(flip_ctx *ctx, int on);
int flip_reset_kernel_timing(flip_ctx *ctx);
int flip_get_kernel_timing(const flip_ctx *ctx, int kernel_class, double *total_ms, int64_t *launches);

/* Number of kernels this library launched on the context since creation. */
int flip_get_kernel_launches(const flip_ctx *ctx, int64_t *launches);
/* Raw cudaStream_t of the context (for callers that time with their own events). */
int flip_get_stream(const flip_ctx *ctx, void **stream);
int flip_synchronize(flip_ctx *ctx);

/* ---- multi-GPU z-slab decomposition ------------------------------------------- */

/* Make this context one z-slab of the global I x J x K domain given to flip_create, shared by the
 * `nranks` processes of one box (one process per GPU).  Rank r owns the global cell planes of
 * flip_slab_range(K, nranks, r) and keeps `halo` (default 16) extra planes towards each neighbour.
 * nccl_unique_id: the 128 bytes of an ncclUniqueId created by rank 0 (flip_get_nccl_unique_id) and
 * distributed by the caller (e.g. a torch.distributed broadcast).  Must precede flip_initialize.
 * Afterwards: flip_load_particles / flip_set_particles may be handed the whole scene (each rank keeps
 * its own planes); flip_get_particles returns the owned particles; flip_get_array / flip_set_array
 * address the LOCAL arrays (flip_get_slab_info); flip_set_solid_sdf still takes the GLOBAL array;
 * flip_step_stats counts are global.  New: the reference has no multi-process path. */
int flip_set_slab(flip_ctx *ctx, int rank, int nranks, const void *nccl_unique_id, int id_bytes);
int flip_get_nccl_unique_id(void *out_id, int id_bytes);
int flip_slab_range(int K, int nranks, int rank, int *k0, int *k1);
int flip_get_slab_info(const flip_ctx *ctx, int *k_offset, int *k_local, int *k_own0, int *k_own1);
int flip_set_halo(flip_ctx *ctx, int planes);

#ifdef __cplusplus
}
#endif
#endif /* FLIP_B200_H */
