# Addition to the plugin's linux/CMakeLists.txt (reference: linux/CMakeLists.txt:46-49, where
# `face_detection_tflite_bundled_libraries` is set to an empty list because flutter_litert bundles the TFLite runtime).
# Bundles the prebuilt CUDA library next to the application so that DynamicLibrary.open('libfdt_cuda.so') finds it.
#
#   include(${CMAKE_CURRENT_SOURCE_DIR}/CMakeLists.fdt.cmake)     # before the PARENT_SCOPE export
#
# Build the library first: python -m face_detection_tflite_b200.build  (nvcc -gencode arch=compute_100a,code=sm_100a)

set(FDT_CUDA_LIBRARY "${CMAKE_CURRENT_SOURCE_DIR}/lib/libfdt_cuda.so" CACHE FILEPATH "prebuilt libfdt_cuda.so (sm_100a)")
if(NOT EXISTS "${FDT_CUDA_LIBRARY}")
  message(FATAL_ERROR "libfdt_cuda.so not found at ${FDT_CUDA_LIBRARY}: build it with `python -m face_detection_tflite_b200.build` "
                      "and copy face_detection_tflite_b200/libfdt_cuda.so to linux/lib/ (there is no CPU fallback)")
endif()

set(face_detection_tflite_bundled_libraries
  "${FDT_CUDA_LIBRARY}"
  PARENT_SCOPE
)
