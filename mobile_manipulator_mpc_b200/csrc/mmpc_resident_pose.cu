// mmpc_resident_pose.cu -- the resident kernel (mmpc_resident.cu) compiled once more with the end-point pose cost of the
// reference's pose-reference controller (controllers/mpc_wholebody.py:49-128; MMPC_MODEL_POSEREF)
// switched in.  Its own translation unit, namespace and entry points, so that the kernels of the whole-body and base-only
// controllers carry none of it: the pose cost (three more gradients and an exact 6 x 6 Hessian per stage) sits behind
// `#ifdef MMPC_POSEREF` in the shared phase bodies (Inst::pose_cost, mmpc_staged.cuh).
#define MMPC_POSEREF 1
#define mmpc_res mmpc_res_pose
#define mmpc_resident_smem_bytes mmpc_resident_pose_smem_bytes
#define mmpc_resident_launch mmpc_resident_pose_launch
#define mmpc_resident_tail_node mmpc_resident_pose_tail_node
#include "mmpc_resident.cu"
