// Packed fixed-point min-sum throughput kernel (IMS_DEC) -- placeholder until the packed kernel lands:
// plan_ims_fast reports "no fast kernel", so IMS_DEC handles run the table-driven kernel.
#include "kernels.h"

namespace ldpcb200 {

FastPlan plan_ims_fast(const QcHost&, const DecParams&, int, int) { return FastPlan(); }

cudaError_t launch_ims_fast(const FastPlan&, const QcDev&, const DecParams&, const FrameIO&, double*, int, cudaStream_t)
{
    return cudaErrorNotSupported;
}

} // namespace ldpcb200
