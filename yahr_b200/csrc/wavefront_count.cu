// wavefront_count.cu -- the COUNTING build of the wavefront kernel set: wavefront.cu compiled with YB_COUNT_WORK, i.e.
// the same kernels in namespace yb::counted with per-lane work counters (see the head of wavefront.cu).  Used by
// yahr_b200_render_device_counted for the `gpu_counted` block of bench.py's roofline; never on the product path.
#define YB_COUNT_WORK 1
#include "wavefront.cu"
