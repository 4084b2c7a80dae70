/* stand-in for the header CMake configures from src/buildConfig.h.in (oracle build only) */
#ifndef EMUO_BUILDCONFIG_H
#define EMUO_BUILDCONFIG_H
#define VERSION_NUMBER "oracle-shim"
#endif
