// Score-history build of the batch-tiled forward kernel: see the header of
// forward_tiled.cu.  Same source, compiled with the argmax bookkeeping, the re-scan
// and the backpointer stores removed; writes the fp32 score history instead.
#define TORBI_TILED_HIST 1
#include "forward_tiled.cu"
