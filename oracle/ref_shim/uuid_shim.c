/* libuuid stand-in for the oracle/_ref build: random v4 UUIDs from /dev/urandom. */
#include <stdio.h>
typedef unsigned char uuid_t[16];
void uuid_generate(uuid_t out){ FILE*f=fopen("/dev/urandom","rb"); if(f){ if(fread(out,1,16,f)!=16){} fclose(f);} out[6]=(out[6]&0x0f)|0x40; out[8]=(out[8]&0x3f)|0x80; }
void uuid_unparse(const uuid_t u, char*o){ sprintf(o,"%02x%02x%02x%02x-%02x%02x-%02x%02x-%02x%02x-%02x%02x%02x%02x%02x%02x",u[0],u[1],u[2],u[3],u[4],u[5],u[6],u[7],u[8],u[9],u[10],u[11],u[12],u[13],u[14],u[15]); }
