#include <osmocom/core/crcgen.h>
