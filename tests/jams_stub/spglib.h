/* tests/jams_stub/spglib.h — TEST INFRASTRUCTURE: the one declaration core/lattice.h needs from spglib (absent here), so that the
 * JAMS-side adapter can be compiled against the reference's real headers (tests/test_adapter_compile.py). */
#ifndef JB_STUB_SPGLIB_H
#define JB_STUB_SPGLIB_H
struct SpglibDataset;
#endif
