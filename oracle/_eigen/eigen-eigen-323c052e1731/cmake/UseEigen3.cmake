#                                               -*- cmake -*-
#
#  UseEigen3.cmake

add_definitions     ( ${EIGEN3_DEFINITIONS} )
include_directories ( ${EIGEN3_INCLUDE_DIRS} )
