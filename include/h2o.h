/*
 * h2o.h -- C ABI of the B200-native hydrodynamics force engine (libh2o_b200.so).
 *
 * Drop-in boundary for ONE path of Joagai23/silver2_isaacsim: the per-body, per-step
 * hydrodynamic force/torque evaluation.  Every entry point names the reference
 * interface it replaces (paths relative to /root/reference/src/scripts/physics/).
 *
 * Conventions
 *   - plain C: opaque handle, raw device pointers + sizes (or DLPack DLTensor*, see
 *     h2o_dlpack.h); no C++/torch types cross the boundary.
 *   - every function returns an h2o_status (0 = ok) and never throws;
 *     h2o_last_error() returns a thread-local message for the last failure.
 *   - all tensors are caller-owned, contiguous row-major, dtype == handle dtype,
 *     on the handle's device, 16-byte aligned.  The engine borrows them for the
 *     duration of the call (or until h2o_unbind for bound tensors) and never frees them.
 *     The engine owns only: coefficient records, previous-step velocities, statistics,
 *     captured graphs and pinned staging buffers.
 *   - every launch takes an explicit cudaStream_t (pass torch.cuda.current_stream().cuda_stream);
 *     there are no hidden synchronisations on the step path.
 *   - a handle is not re-entrant: one caller thread at a time
 *     (the reference is called serially from the PhysX step callback,
 *     hydrodynamics_behavior.py:131-141).
 */
#ifndef H2O_H_
#define H2O_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define H2O_API __attribute__((visibility("default")))
#else
#define H2O_API
#endif

typedef struct h2o_engine* h2o_handle;
typedef void* h2o_stream; /* cudaStream_t */

typedef enum {
    H2O_OK = 0,
    H2O_ERR_BAD_HANDLE = 1,
    H2O_ERR_BAD_ARGUMENT = 2,
    H2O_ERR_BAD_SHAPE = 3,
    H2O_ERR_BAD_DTYPE = 4,
    H2O_ERR_BAD_DEVICE = 5,
    H2O_ERR_NOT_CONTIGUOUS = 6,
    H2O_ERR_ALIGNMENT = 7,
    H2O_ERR_NOT_CONFIGURED = 8, /* parameters / bound tensors / rollout missing */
    H2O_ERR_CUDA = 9,
    H2O_ERR_NO_DEVICE = 10
} h2o_status;

typedef enum { H2O_F32 = 0, H2O_F64 = 1 } h2o_dtype;
typedef enum { H2O_QUAT_XYZW = 0, H2O_QUAT_WXYZ = 1 } h2o_quat_order;
typedef enum { H2O_KERNEL_AUTO = 0, H2O_KERNEL_TILE = 1, H2O_KERNEL_DIRECT = 2 } h2o_kernel_choice;

/* Number of scalars in one coefficient record and their order:
 * xDimension, yDimension, zDimension, linearDragCoefficient, angularDragCoefficient,
 * linearDamping, angularDamping, linearAddedMassCoefficient, angularAddedMassCoefficient,
 * liftCoefficient, mass      (names: hydrodynamics_behavior.py:28-46; mass: :172-173) */
#define H2O_N_COEFF 11
/* h2o_stats vector: sum|F|, max|F|, wet bodies, clamped bodies, non-finite forces,
 * wet-and-at-rest bodies (the reference raises there, numba_hydrodynamics.py:118,143),
 * bodies processed, bodies the fp32 fast path handed to the float64 re-evaluation. */
#define H2O_N_STATS 8

H2O_API const char* h2o_last_error(void);
H2O_API const char* h2o_version(void);
/* Number of CUDA devices visible, or a negative h2o_status. */
H2O_API int h2o_device_count(void);

/* ---- construction -------------------------------------------------------------------
 * Replaces XHydrodynamicsWrapper.__init__ (numba_hydrodynamics_wrapper.py:9-32,
 * warp_hydrodynamics_wrapper.py:10-77), batched over n_bodies, and
 * HydrodynamicsBehavior._setup (hydrodynamics_behavior.py:143-174). */
H2O_API int h2o_create(h2o_handle* out, int64_t n_bodies, int dtype, int device);
H2O_API int h2o_destroy(h2o_handle h);

/* waterDensity, gravity (hydrodynamics_behavior.py:30-31; hydrodynamics_config.json "globals") */
H2O_API int h2o_set_globals(h2o_handle h, double water_density, double gravity);

/* Non-flat water surface: eta is a DEVICE array (n_bodies,) in the handle's dtype
 * holding the surface elevation above surface_z at each body's position (e.g. a wave field sampled
 * by the caller); the model then sees p_z - surface_z - eta[i] where the reference uses p_z
 * (analyze_submersion_and_cob, numba_hydrodynamics.py:59-105, tests keypoints against z = 0).  The
 * pointer is borrowed and read by every later step until replaced (update the array in place between
 * steps; replacing the pointer drops a rollout captured with h2o_capture_rollout); NULL =
 * flat.  Steps with a height field run on the per-body kernel. */
H2O_API int h2o_set_surface_heights(h2o_handle h, const void* eta_dev);

/* Dense added mass.  calculate_added_mass (numba_hydrodynamics.py:219-253) takes
 * ANY 6x6 body-frame matrix M: f6 = -M [R^T a; R^T alpha], F = R f6[0:3] ratio, tau = R f6[3:6] ratio;
 * the wrapper only ever builds a diagonal one (numba_hydrodynamics_wrapper.py:101-112).  This call
 * installs n_types row-major 6x6 matrices (host, float64) and a slot -> matrix map: body i uses
 * matrices[slot_type[i % n_slots]] INSTEAD of the C_a / C_a_omega diagonal of its coefficient record.
 * slot_type may be NULL when n_types == 1.  n_types == 0 restores the diagonal.  Steps with dense
 * matrices run on the per-body kernel (H2O_KERNEL_DIRECT); h2o_components keeps the diagonal. */
H2O_API int h2o_set_added_mass_dense(h2o_handle h, int n_types, const double* matrices, int n_slots,
                                     const int32_t* slot_type);

/* Generalisation the reference's signatures admit but its scenes never use:
 * a uniform water current (drag, damping and lift then see v - current) and the height of the flat
 * water surface (the reference's is the plane z = 0).  NULL / 0 = the reference's behaviour. */
H2O_API int h2o_set_environment(h2o_handle h, const double current_xyz[3], double surface_z);

/* Same twelve scalars, same order as the reference wrapper ctor
 * (numba_hydrodynamics_wrapper.py:9-10): width, depth, height, linear_drag_coefficient,
 * angular_drag_coefficient, linear_damping, angular_damping, water_density, gravity,
 * linear_mass_coeff, angular_mass_coeff, lift_coefficient; plus the body mass used by the
 * safety clamp (hydrodynamics_behavior.py:172-173, :221-226).  Applies to every body. */
H2O_API int h2o_set_params_uniform(h2o_handle h, const double ctor12[12], double mass);

/* Part-type table (hydrodynamics_config.json "parts"): table is n_types x H2O_N_COEFF host
 * doubles; slot_type[n_slots] maps body slot (global body index mod n_slots) to a type.
 * The table is staged in shared memory by the kernels. */
H2O_API int h2o_set_part_table(h2o_handle h, int n_types, const double* table_host, int n_slots,
                               const int32_t* slot_type_host);

/* Heterogeneous per-body records: coeff is (n_bodies, H2O_N_COEFF), host or device memory,
 * src_dtype H2O_F32/H2O_F64 (converted to the handle dtype on the device). */
H2O_API int h2o_set_params_per_body(h2o_handle h, const void* coeff, int src_dtype, h2o_stream stream);

/* Same, struct-of-arrays: cols[k] is a DEVICE array (n_bodies,) holding coefficient k of every body, in
 * H2O_N_COEFF order (xDimension, yDimension, zDimension, linearDragCoefficient, angularDragCoefficient,
 * linearDamping, angularDamping, linearAddedMassCoefficient, angularAddedMassCoefficient,
 * liftCoefficient -- the per-prim exposed variables of hydrodynamics_behavior.py:32-44 -- and mass, the
 * one _setup reads from the view, :172-174).  Interleaved into the engine's records on the device. */
H2O_API int h2o_set_params_soa(h2o_handle h, const void* const cols[11], int src_dtype, h2o_stream stream);

/* Articulation structure: bodies are grouped in contiguous runs of bodies_per_robot;
 * 0 disables the per-robot wrench. */
H2O_API int h2o_set_articulation(h2o_handle h, int bodies_per_robot);

/* Articulation of UNEQUAL robots (e.g. the reference's main scene: one 19-link SILVER2 plus the Obsea
 * buoy as a robot of its own): robot r owns bodies [offsets[r], offsets[r+1]); offsets is a host array
 * of n_robots + 1 entries, offsets[0] = 0, offsets[n_robots] = n_bodies, strictly increasing.  The
 * out_robot_wrench of the step calls is then (n_robots, 6), each wrench about its robot's first body,
 * reduced from the written force / torque arrays after the force kernel.  h2o_set_articulation
 * replaces it again.  Not available through h2o_step_host. */
H2O_API int h2o_set_articulation_offsets(h2o_handle h, int64_t n_robots, const int64_t* offsets);

/* Isaac core hands quaternions as wxyz and the reference permutes them
 * (hydrodynamics_behavior.py:194); the wrappers themselves take xyzw. */
H2O_API int h2o_set_quat_order(h2o_handle h, int order);
H2O_API int h2o_set_kernel(h2o_handle h, int choice);
/* Follow the reference's WARP twin instead of its Numba path (warp_hydrodynamics.py; SURVEY.md Appendix C;
 * pinned by tests/golden/reference_warp_golden.npz, vectors of the reference's own kernel source): accelerations
 * rotated forward instead of inverse for the added-mass terms (:216-217), every rotation by wp.quat_rotate
 * (= R(q) + 2(|q|^2 - 1) I for an un-normalised quaternion), centre of buoyancy = mean of the wet keypoints also
 * when fully submerged (:57-58), cob = cop = position for a dry body (:59-61, :290).  Applies to h2o_components
 * and to the fused step entry points (the production behaviour script evaluates forces through the Warp wrapper,
 * hydrodynamics_behavior.py:19, :155); the step then runs every body through the float64 formulation on the
 * per-body kernel (a compatibility mode, not a fast path).  Default 0 = the Numba semantics (north star). */
H2O_API int h2o_set_warp_compat(h2o_handle h, int enable);
/* fp32 mode only.  1 (default): every body meets the fp32-mode bound of the parity criterion (|dF|, |dtau| <=
 * max(1e-5 |.|_inf, 1e-6) against the float64 Numba path): the fused step flags the few bodies per 10 000 whose
 * force / torque groups cancel (or whose quaternion is far from unit) and re-evaluates them in float64 with the
 * world-frame formulation of solve_hydrodynamics (numba_hydrodynamics.py:255-314).  0: flagged bodies keep their
 * fp32 result (about one body per million then misses the bound by up to 3x); saves the re-evaluation tail. */
H2O_API int h2o_set_strict(h2o_handle h, int enable);
/* Tuning knob: tile-kernel variant (threads per CTA / TMA stages); 0 = built-in default. */
H2O_API int h2o_set_tile_config(h2o_handle h, int cfg);
/* Accumulate global statistics inside the step kernel (device-side, no host sync). */
H2O_API int h2o_enable_stats(h2o_handle h, int enable);

/* ---- carried state ------------------------------------------------------------------
 * _last_linear_velocity / _last_angular_velocity (hydrodynamics_behavior.py:196-198,
 * :237-238).  h2o_reset mirrors _reset (:240-245): the next step sees zeros. */
H2O_API int h2o_reset(h2o_handle h, h2o_stream stream);
H2O_API int h2o_set_prev(h2o_handle h, const void* prev_lin, const void* prev_ang, h2o_stream stream);
H2O_API int h2o_get_prev(h2o_handle h, void* prev_lin, void* prev_ang, h2o_stream stream);

/* ---- fused step ---------------------------------------------------------------------
 * Replaces HydrodynamicsBehavior._apply_behavior (hydrodynamics_behavior.py:194-238) for all
 * bodies at once: quaternion reorder, finite-difference acceleration, every force term,
 * body-relative lever arms, net wrench, safety clamp, previous-velocity update.
 *   pos (N,3) quat (N,4) lin_vel (N,3) ang_vel (N,3)  ->  out_force (N,3) out_torque (N,3)
 *   out_robot_wrench: (N / bodies_per_robot, 6) [F, tau about the robot's slot-0 body] or NULL.
 * dt <= 1e-6 is a successful no-op (hydrodynamics_behavior.py:139). */
H2O_API int h2o_step(h2o_handle h, const void* pos, const void* quat, const void* lin_vel,
                     const void* ang_vel, double dt, void* out_force, void* out_torque,
                     void* out_robot_wrench, h2o_stream stream);

/* Same, PhysX tensor-API layout: transforms (N,7) = [p, q], velocities (N,6) = [v, w]
 * (RigidPrimView.get_world_poses / get_velocities, hydrodynamics_behavior.py:178-189). */
H2O_API int h2o_step_physx(h2o_handle h, const void* transforms, const void* velocities, double dt,
                           void* out_force, void* out_torque, void* out_robot_wrench,
                           h2o_stream stream);

/* Same, RigidPrimView layout: pos (N,3), quat (N,4) from get_world_poses and the fused
 * velocities (N,6) = [v, w] from get_velocities (hydrodynamics_behavior.py:178-189) -- no slicing
 * copies on the caller's side. */
H2O_API int h2o_step_view(h2o_handle h, const void* pos, const void* quat, const void* velocities, double dt,
                          void* out_force, void* out_torque, void* out_robot_wrench, h2o_stream stream);

/* Bind the tensors once, then step with a single cheap call (layout: 0 split, 1 physx, 2 view;
 * physx: pass transforms as pos, velocities as lin_vel, quat = ang_vel = NULL;
 * view: pass velocities as lin_vel, ang_vel = NULL). */
H2O_API int h2o_bind(h2o_handle h, int layout, const void* pos, const void* quat, const void* lin_vel,
                     const void* ang_vel, void* out_force, void* out_torque, void* out_robot_wrench);
H2O_API int h2o_unbind(h2o_handle h);
H2O_API int h2o_step_bound(h2o_handle h, double dt, h2o_stream stream);

/* CUDA-graph rollout over the bound tensors: n_steps back-to-back steps captured once,
 * replayed with one launch (the reference captures a single dim=1 kernel,
 * warp_hydrodynamics_wrapper.py:101-120).  Capturing launches nothing: the carried velocities and
 * (free-body mode) the bound state are exactly as before the call.  The graph bakes in the engine's
 * device buffers and constants, so every h2o_set_* / h2o_enable_stats / h2o_bind / h2o_unbind call drops
 * it: h2o_launch_rollout then returns H2O_ERR_NOT_CONFIGURED until the rollout is captured again. */
H2O_API int h2o_capture_rollout(h2o_handle h, int n_steps, double dt, h2o_stream stream);
/* Rollout mode: 0 = static state (force-only rollout), 1 = free bodies: after each step the bound
 * pose / velocity tensors are integrated in place by h2o_integrate_free_bodies (split layout). */
H2O_API int h2o_set_rollout_mode(h2o_handle h, int free_bodies, double gravity);
H2O_API int h2o_launch_rollout(h2o_handle h, h2o_stream stream);

/* Persistent rollout of FREE bodies over the bound tensors (split layout): n_steps x (fused force step ->
 * h2o_integrate_free_bodies) in ONE kernel launch, each thread carrying its body's pose, velocities,
 * previous velocities and coefficients in registers across the steps (bodies are independent:
 * solve_hydrodynamics reads only its own body, numba_hydrodynamics.py:255-314; the behaviour's per-step
 * bookkeeping is hydrodynamics_behavior.py:196-238).  Same result as h2o_capture_rollout in free-body mode
 * (up to FMA contraction) without a launch / graph node per step.  On return of the kernel the bound pose /
 * velocity tensors and the carried velocities hold the final state, out_force / out_torque the last step's
 * wrench.  trace_every > 0: every trace_every-th step a row [p, v, w] (9 scalars, handle dtype) per body
 * goes to trace_out, a device array (n_steps / trace_every, n_bodies, 9) -- the columns of the
 * reference's velocity logger (log_velocity.py:17-20).  dt <= 1e-6 is a successful no-op. */
H2O_API int h2o_rollout_persistent(h2o_handle h, int n_steps, double dt, double gravity, int trace_every,
                                   void* trace_out, h2o_stream stream);

/* ---- stand-alone free-body stepper (harness; the reference leaves integration to PhysX) ------
 * Semi-implicit Euler for free boxes (mass and dimensions from the coefficient records, box
 * inertia, gravity along -z): velocities first, then pose; all four state tensors updated in place. */
H2O_API int h2o_integrate_free_bodies(h2o_handle h, void* pos, void* quat, void* lin_vel, void* ang_vel,
                                      const void* force, const void* torque, double dt, double gravity,
                                      h2o_stream stream);

/* ---- full-signature components --------------------------------------------------------
 * Replaces XHydrodynamicsWrapper.calculate_hydrodynamic_forces
 * (numba_hydrodynamics_wrapper.py:34-53, warp_hydrodynamics_wrapper.py:79-132) =
 * solve_hydrodynamics (numba_hydrodynamics.py:255-314), batched.  Outputs in the reference's
 * order: buoyancy_force, drag_force, lift_force, drag_torque, added_mass_force,
 * added_mass_torque, center_of_buoyancy, center_of_pressure (each (N,3)), sub_ratio (N,).
 * out_flags (N, int32, may be NULL): bit0 set where the unmodified reference raises TypeError
 * (wet body with speed <= 1e-6); the engine returns cop = cob, area = 0 there. */
H2O_API int h2o_components(h2o_handle h, const void* pos, const void* quat, const void* lin_vel,
                           const void* ang_vel, const void* lin_acc, const void* ang_acc,
                           void* const out8[8], void* out_sub_ratio, int32_t* out_flags,
                           h2o_stream stream);

/* ---- host-buffer path --------------------------------------------------------------------
 * Numba-wrapper style call with HOST arrays (the reference's CPU flavour takes and returns NumPy arrays,
 * numba_hydrodynamics_wrapper.py:34-53).  Synchronous: ordered after the work queued before on the
 * blocking streams (an event on the legacy default stream, no device-wide synchronisation), returns when
 * the results are in the host buffers.
 *   - pinned (page-locked) host buffers, 16-byte aligned: ZERO-COPY.  The fused tile kernel runs directly on
 *     them: its TMA bulk loads pull each tile over PCIe into shared memory, its bulk stores push force /
 *     torque back -- one kernel, transfers overlapped with the arithmetic tile by tile, no staging copy.
 *   - pageable buffers: staged through engine-owned pinned memory, chunked H2D -> step -> D2H pipeline on
 *     three streams.
 * h2o_last_host_path reports which one the last call took (1 zero-copy, 2 staged). */
H2O_API int h2o_step_host(h2o_handle h, const void* pos, const void* quat, const void* lin_vel,
                          const void* ang_vel, double dt, void* out_force, void* out_torque,
                          void* out_robot_wrench);
/* Same with the PhysX tensor-API layout on the host: transforms (N,7) = [p, q], velocities (N,6) = [v, w]
 * (what RigidPrimView.get_world_poses / get_velocities hand out when the simulation runs on the CPU,
 * hydrodynamics_behavior.py:178-189): two input arrays instead of four. */
H2O_API int h2o_step_host_physx(h2o_handle h, const void* transforms, const void* velocities, double dt,
                                void* out_force, void* out_torque, void* out_robot_wrench);
H2O_API int h2o_last_host_path(h2o_handle h);

/* ---- statistics / introspection ----------------------------------------------------------- */
H2O_API int h2o_stats_device_ptr(h2o_handle h, void** out_ptr); /* H2O_N_STATS doubles on device */
H2O_API int h2o_read_stats(h2o_handle h, double out[H2O_N_STATS], int reset, h2o_stream stream);
H2O_API int64_t h2o_launch_count(h2o_handle h); /* kernels launched (or graph nodes replayed) so far */
H2O_API int64_t h2o_n_bodies(h2o_handle h);
H2O_API int h2o_dtype_of(h2o_handle h);
/* Which kernel the last step used: H2O_KERNEL_TILE or H2O_KERNEL_DIRECT. */
H2O_API int h2o_last_kernel(h2o_handle h);
/* Resident CTAs per SM of the last tile-kernel launch (occupancy query result). */
H2O_API int h2o_last_ctas_per_sm(h2o_handle h);
/* Device pointers of engine-owned buffers (zero-copy views for DLPack export on the host side). */
H2O_API int h2o_prev_device_ptr(h2o_handle h, void** out_ptr);  /* (N,6) */
H2O_API int h2o_coeff_device_ptr(h2o_handle h, void** out_ptr, int64_t* out_rows); /* (rows,11) */

#ifdef __cplusplus
}
#endif
#endif /* H2O_H_ */
