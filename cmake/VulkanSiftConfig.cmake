# find_package(VulkanSift) for the B200 build: same variables as the reference's installed package
# (README.md:72-88 of maelaubert/VulkanSift, cmake/VulkanSiftConfig.cmake.in):
#
#   set(VulkanSift_DIR <this repository>/cmake)
#   find_package(VulkanSift)
#   target_link_libraries(TARGET ${VulkanSift_LIB})
#   target_include_directories(TARGET PRIVATE ${VulkanSift_INCLUDE_DIR})
#
# The library is built in-tree by `python -m vulkansift_b200.build` (nvcc, sm_100a); it links cudart statically,
# so consumers need no CUDA toolkit, only the NVIDIA driver at run time.
get_filename_component(_vksift_root "${CMAKE_CURRENT_LIST_DIR}/.." ABSOLUTE)
set(VulkanSift_INCLUDE_DIR "${_vksift_root}/include")
set(_vksift_lib "${_vksift_root}/vulkansift_b200/lib/libvulkansift.so")
if(NOT EXISTS "${_vksift_lib}")
  message(FATAL_ERROR "VulkanSift (B200 build): ${_vksift_lib} not found, run `python -m vulkansift_b200.build` first")
endif()
if(NOT TARGET vulkansift)
  add_library(vulkansift SHARED IMPORTED)
  set_target_properties(vulkansift PROPERTIES
    IMPORTED_LOCATION "${_vksift_lib}"
    INTERFACE_INCLUDE_DIRECTORIES "${VulkanSift_INCLUDE_DIR}")
endif()
set(VulkanSift_LIB vulkansift)
set(VulkanSift_FOUND TRUE)
