#include "../mjp_variant.cu"
