/* libuuid stand-in for the oracle/_ref build (headers absent in this image). */
#pragma once
#ifdef __cplusplus
extern "C" {
#endif
typedef unsigned char uuid_t[16];
void uuid_generate(uuid_t out);
void uuid_unparse(const uuid_t uu, char* out);
#ifdef __cplusplus
}
#endif
