! Stand-in for src/gributils/class_gribfile_mod.f90 (which wraps the eccodes API and cannot be built
! here): only the two centre constants convmix / calcmatrix test against.  Values as in the reference
! (src/gributils/class_gribfile_mod.f90:45-47).  Written for the transpile recipe; test infrastructure.
module class_gribfile
  implicit none
  integer, parameter :: GRIBFILE_CENTRE_NCEP = 1
  integer, parameter :: GRIBFILE_CENTRE_ECMWF = 2
end module class_gribfile
