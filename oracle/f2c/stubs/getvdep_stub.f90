! Stand-in for src/getvdep.f90 (dry-deposition velocities from land use, season and surface
! resistances: getrb, raerod, getrc, partdep and the landuse inventory it needs are outside the path
! this repo covers).  calcpar only calls it when DRYDEP is set; the transpile recipe needs the symbol.
! Written for the transpile recipe; test infrastructure.
subroutine getvdep(n,ix,jy,ust,temp,pa,L,gr,rh,rr,snow,vdepo)
  use par_mod
  implicit none
  integer :: n,ix,jy,i
  real :: ust,temp,pa,L,gr,rh,rr,snow
  real :: vdepo(maxspec)
  do i=1,maxspec
    vdepo(i)=0.
  end do
end subroutine getvdep
