// oracle build only: minimal stand-in for the CMake-generated Xyce_config.h
// (template: /root/reference/src/Xyce_config.h.cmake).  No optional feature is enabled.
#ifndef Xyce_CONFIG_H
#define Xyce_CONFIG_H
#define HAVE_UNISTD_H
#endif
