#ifndef COMMON_CONFIG_SHIM_H
#define COMMON_CONFIG_SHIM_H
/* what CMake's configure_file would emit with every optional target off */
#endif
