"""hippopt_b200 -- B200-native evaluation path for hippopt's multiple-shooting planners.

Modules (nothing is imported eagerly: `_capi.lib()` loads libhippopt_b200.so on first use and raises if it is
missing -- there is no CPU fallback, and nothing here imports `oracle/`):

  _capi            ctypes binding of include/hippopt_b200.h (enums parsed from the header)
  robot_model      kinematic tree from a URDF string (adam's lumping of unlisted joints) / the synthetic ergoCub
  kino_layout      x / p / g ordering and CCS patterns of the kinodynamic OCP, scatter maps for the kernels
  pose_layout      the same for the pose finder
  evaluator        KinoEvaluator / PoseEvaluator / ToyEvaluator (hb_eval), HostPipeline (hb_eval_host)
  sharding         instance sharding over ranks, gather of per-instance results (torch.distributed)
  workloads        synthetic batches of BASELINE's configurations
  kkt              stage-wise KKT sweep on the batched LU kernels (row f2)
  ipsolver         batched interior-point driver with IPOPT's termination options (row f1)
  opti_callback    the reference's callback criteria, one state per instance (row f1)
  interpolators    humanoid_state_interpolator: schedule on the host, one kernel on the device (row f3)
  initial_guess    batched set-up of periodic-step plans (row f3)
"""
