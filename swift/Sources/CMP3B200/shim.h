/* The C ABI of the B200 encode path (include/mp3b200.h at the repository root; pass -Xcc -I<repo>/include). */
#include "mp3b200.h"
