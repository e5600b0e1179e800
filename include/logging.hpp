/** logging.hpp -- the reference's three logging macros (src/logging.hpp:1-4), kept for source compatibility. */
#ifndef LOGGING
#define LOGGING 0
#define INFO_LOG "[INFO] "
#define DEBUG 0
#endif
