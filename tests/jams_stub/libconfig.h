/* tests/jams_stub/libconfig.h — TEST INFRASTRUCTURE: helpers/error.h includes the C header of libconfig (absent here) and uses
 * nothing from it. */
#ifndef JB_STUB_LIBCONFIG_H
#define JB_STUB_LIBCONFIG_H
#endif
