// The pipelined sweep kernel with its phase timers compiled in (see sweep3.cu); launched when qio_b200_sweep_options.timers is set.
#define QIO_S3_TIMED 1
#define QIO_LS_TIMED 1
#define S3_NS s3t
#define S3_LAUNCH sweep3_timed_launch
#include "sweep3.cu"
