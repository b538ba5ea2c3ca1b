/*
 * h2o_dlpack.h -- DLPack-typed entry points of libh2o_b200.so.
 *
 * The host side (Python/PyTorch) hands tensors over zero-copy as DLPack DLTensor
 * descriptors (torch.utils.dlpack.to_dlpack -> PyCapsule "dltensor" ->
 * DLManagedTensor*), the same way the reference's GPU wrapper wraps torch memory with
 * wp.from_torch (warp_hydrodynamics_wrapper.py:85-90) -- but without the six staging
 * copies that follow there (:93-98).  The engine validates dtype / device / shape /
 * contiguity / alignment from the descriptor, borrows the memory for the call and never
 * takes ownership (the capsule stays with the caller).
 *
 * The struct declarations below restate the public DLPack ABI (dmlc/dlpack v0.8,
 * DLManagedTensor flavour) so that this header is self-contained plain C; they are
 * skipped when the official <dlpack/dlpack.h> was included first.
 */
#ifndef H2O_DLPACK_H_
#define H2O_DLPACK_H_

#include <stdint.h>

#include "h2o.h"

#ifdef __cplusplus
extern "C" {
#endif

#ifndef DLPACK_DLPACK_H_
#define DLPACK_DLPACK_H_
typedef enum {
    kDLCPU = 1,
    kDLCUDA = 2,
    kDLCUDAHost = 3,
    kDLCUDAManaged = 13
} DLDeviceType;
typedef struct {
    DLDeviceType device_type;
    int32_t device_id;
} DLDevice;
typedef enum { kDLInt = 0U, kDLUInt = 1U, kDLFloat = 2U, kDLBfloat = 4U } DLDataTypeCode;
typedef struct {
    uint8_t code;
    uint8_t bits;
    uint16_t lanes;
} DLDataType;
typedef struct {
    void* data;
    DLDevice device;
    int32_t ndim;
    DLDataType dtype;
    int64_t* shape;
    int64_t* strides; /* in elements; NULL = compact row-major */
    uint64_t byte_offset;
} DLTensor;
typedef struct DLManagedTensor {
    DLTensor dl_tensor;
    void* manager_ctx;
    void (*deleter)(struct DLManagedTensor* self);
} DLManagedTensor;
#endif /* DLPACK_DLPACK_H_ */

/* DLPack twins of h2o_step / h2o_step_physx / h2o_bind / h2o_components /
 * h2o_set_params_per_body (see h2o.h for the reference interfaces they replace).
 * Shapes: (N,3) / (N,4) / (N,6) / (N,7) / (N,11) / (N,) with N == h2o_n_bodies(h);
 * out_robot_wrench (N / bodies_per_robot, 6) may be NULL. */
H2O_API int h2o_step_dl(h2o_handle h, const DLTensor* pos, const DLTensor* quat,
                        const DLTensor* lin_vel, const DLTensor* ang_vel, double dt,
                        const DLTensor* out_force, const DLTensor* out_torque,
                        const DLTensor* out_robot_wrench, h2o_stream stream);
H2O_API int h2o_step_physx_dl(h2o_handle h, const DLTensor* transforms, const DLTensor* velocities,
                              double dt, const DLTensor* out_force, const DLTensor* out_torque,
                              const DLTensor* out_robot_wrench, h2o_stream stream);
H2O_API int h2o_step_view_dl(h2o_handle h, const DLTensor* pos, const DLTensor* quat, const DLTensor* velocities,
                             double dt, const DLTensor* out_force, const DLTensor* out_torque,
                             const DLTensor* out_robot_wrench, h2o_stream stream);
H2O_API int h2o_bind_dl(h2o_handle h, int layout, const DLTensor* pos, const DLTensor* quat,
                        const DLTensor* lin_vel, const DLTensor* ang_vel, const DLTensor* out_force,
                        const DLTensor* out_torque, const DLTensor* out_robot_wrench);
H2O_API int h2o_components_dl(h2o_handle h, const DLTensor* pos, const DLTensor* quat,
                              const DLTensor* lin_vel, const DLTensor* ang_vel,
                              const DLTensor* lin_acc, const DLTensor* ang_acc,
                              const DLTensor* const out8[8], const DLTensor* out_sub_ratio,
                              const DLTensor* out_flags, h2o_stream stream);
H2O_API int h2o_set_params_per_body_dl(h2o_handle h, const DLTensor* coeff, h2o_stream stream);
H2O_API int h2o_set_prev_dl(h2o_handle h, const DLTensor* prev_lin, const DLTensor* prev_ang,
                            h2o_stream stream);

/* Export the engine-owned previous-velocity buffer (N,6) as a DLManagedTensor (zero-copy;
 * the deleter frees only the descriptor -- the memory lives as long as the handle). */
H2O_API int h2o_export_prev_dl(h2o_handle h, DLManagedTensor** out);

#ifdef __cplusplus
}
#endif
#endif /* H2O_DLPACK_H_ */
