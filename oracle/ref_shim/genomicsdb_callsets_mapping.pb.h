#pragma once
class CallsetMappingPB;
