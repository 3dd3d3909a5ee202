"""Location of the compiled extension, mirroring the reference's `clode.cpp.clode_cpp_wrapper`."""
