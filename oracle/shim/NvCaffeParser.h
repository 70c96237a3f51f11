// oracle/shim/NvCaffeParser.h -- compile-only stand-in.  TEST INFRASTRUCTURE.
#pragma once
namespace nvcaffeparser1 { class IPluginFactory; }
