# lidar_processing_libConfig.cmake - lets the reference's ROS2 node find the B200 library under the name it
# already asks for (src/processor/CMakeLists.txt:17 `find_package(lidar_processing_lib REQUIRED)`, :51-56
# `target_link_libraries(processor lidar_processing_lib ...)`), replacing the package the reference installs from
# lidar_processing_lib/CMakeLists.txt:89-125 (lidar_processing_libConfig.cmake + lidar_processing_libTargets.cmake).
#
#   colcon build --cmake-args -Dlidar_processing_lib_DIR=/opt/lidar-b200/cmake
#
# No edit of the node: the imported target carries the adaptor headers (include/lidar_processing_lib/*.hpp, same
# class / enum / struct names) and links liblpl_b200.so.
get_filename_component(LPL_B200_ROOT "${CMAKE_CURRENT_LIST_DIR}/.." ABSOLUTE)
set(LPL_B200_LIBRARY "${LPL_B200_ROOT}/lidar_processing_v2_b200/liblpl_b200.so")

if(NOT EXISTS "${LPL_B200_LIBRARY}")
  set(lidar_processing_lib_FOUND FALSE)
  set(lidar_processing_lib_NOT_FOUND_MESSAGE
      "${LPL_B200_LIBRARY} is missing: build it with `python -m lidar_processing_v2_b200.build` (nvcc, sm_100a); there is no CPU fallback")
  return()
endif()

if(NOT TARGET lidar_processing_lib)
  add_library(lidar_processing_lib SHARED IMPORTED GLOBAL)
  set_target_properties(lidar_processing_lib PROPERTIES
    IMPORTED_LOCATION "${LPL_B200_LIBRARY}"
    IMPORTED_NO_SONAME TRUE
    INTERFACE_INCLUDE_DIRECTORIES "${LPL_B200_ROOT}/include"
    INTERFACE_COMPILE_FEATURES cxx_std_17)
endif()

set(lidar_processing_lib_INCLUDE_DIRS "${LPL_B200_ROOT}/include")
set(lidar_processing_lib_LIBRARIES lidar_processing_lib)
set(lidar_processing_lib_FOUND TRUE)
