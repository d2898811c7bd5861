#ifndef PHYSIS_CONFIG_SHIM_H
#define PHYSIS_CONFIG_SHIM_H
/* what CMake's configure_file would emit with every optional target off */
#endif
