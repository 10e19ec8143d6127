#ifdef __cplusplus
#include <cstdint>
#include <cstring>
#include <string>
#endif
#ifndef GENOMICSDB_VERSION
#define GENOMICSDB_VERSION "0.10.2-oracle"
#endif
