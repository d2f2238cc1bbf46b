// measurement build of the default screen: the product source with clock64 stamps compiled in (tools/screen_trace.py)
#define TSC_SCREEN_TRACE 1
#include "../../tscode_b200/csrc/rmsd_screen.cu"
