// The binary32-state batch kernels as their own translation unit: same source as
// smart_kernels.cu, compiled WITHOUT -fmad=false (see the note at the top of that file).
// Emits smart_batch_run_f32 only.
#define SMART_TU_F32
#include "smart_kernels.cu"
