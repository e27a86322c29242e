// pg_ctc_fused_kernel<16, 512, *, *>: 16 CTC states per lane (see fused_impl.cuh)
#include "fused_impl.cuh"

namespace pgasr {
int launch_fused_spl16(int mode, FusedArgs& a, size_t smem, cudaStream_t st) {
    return launch_fused_modes<16, 512>(mode, a, smem, st);
}
}  // namespace pgasr
