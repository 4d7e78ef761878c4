#include "../../include/sgqn_b200.h"
extern "C" int sgqn_abi_version(void) { return 2; }
