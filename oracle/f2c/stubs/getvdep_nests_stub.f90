! Stand-in for src/getvdep_nests.f90 (see getvdep_stub.f90): calcpar_nests only calls it when DRYDEP is
! set; the transpile recipe needs the symbol.  Written for the transpile recipe; test infrastructure.
subroutine getvdep_nests(n,ix,jy,ust,temp,pa,L,gr,rh,rr,snow,vdepo,lnest)
  use par_mod
  implicit none
  integer :: n,ix,jy,i,lnest
  real :: ust,temp,pa,L,gr,rh,rr,snow
  real :: vdepo(maxspec)
  do i=1,maxspec
    vdepo(i)=0.
  end do
end subroutine getvdep_nests
