#pragma once
#include <string>
namespace google { namespace protobuf { inline void ShutdownProtobufLibrary() {} } }
class VidMappingPB;
