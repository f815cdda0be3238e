// tsdf_host.h -- host-side helpers shared by the engine, the C ABI and the CLI.
#pragma once
// Sets the calling thread's tsdf_last_error() text (defined in tsdf_engine.cu).
void tsdf_set_error_text(const char* text);
