// hippopt_b200.cu -- single translation unit of libhippopt_b200.so (kernels + C ABI).
#include "kino_kin.cu"
#include "kino_contact.cu"
#include "pose_contact.cu"
#include "toy.cu"
#include "lu.cu"
#include "kkt_assemble.cu"
#include "interp.cu"
#include "api.cu"
