// TEST INFRASTRUCTURE: see standin_blocks.h
#include "standin_blocks.h"
