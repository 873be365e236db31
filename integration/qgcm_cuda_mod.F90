! qgcm_cuda_mod.F90 -- the Fortran side of the drop-in: everything the reference's main program
! (src/q-gcm.F) needs to run its time loop on libqgcm_b200.so.  Add this file and
! integration/qgcm_types.f90 (generated from include/qgcm_b200.h) to src/Makefile's module list
! after intrfac_data.F / ochomog_data.F / athomog_data.F / monitor_data.F, and link with
! -lqgcm_b200 -lcudart.  The three insertions in q-gcm.F are listed in INTEGRATION.md section 3.
!
! Free-form Fortran 2003 (ISO_C_BINDING); compile with the reference's cpp macros
! (-Docean_only, -Dcyclic_ocean, ...), exactly like the other .F files.
!
! Status: written against the C header and the reference's module declarations; NOT compiled
! (no Fortran compiler exists in the build image or on the GPU box, profiles/r02_fortran_probe.txt).
! tests/test_integration_glue.py checks statically that every qgcm_* entry point used here is
! declared in the header with that many arguments, that every field name passed to
! qgcm_set_field / qgcm_get_field is a field the library registers, and that every module
! variable named here is declared in the reference's *_data.F modules.
module qgcm_cuda
  use, intrinsic :: iso_c_binding
  use qgcm_types
  implicit none
  private
  public :: gpu, gpu_init, gpu_finalize, gpu_check, gpu_upload_state, gpu_download, gpu_sync_scalars
  public :: GPU_FOR_VALIDS, GPU_FOR_MONNC, GPU_FOR_OCNC, GPU_FOR_ATNC, GPU_FOR_TAVOCN, GPU_FOR_TAVATM, &
            GPU_FOR_QOCDIAG, GPU_FOR_RESTART, GPU_FOR_PRSAMP, GPU_FOR_AREAVG, GPU_FOR_COVOCN, GPU_FOR_COVATM

  type(c_ptr), save :: gpu = c_null_ptr          ! the opaque qgcm_model*

  ! qgcm_host_register with the array itself as argument (the generated interface takes a C
  ! address, and c_loc would need the TARGET attribute the reference's arrays do not have)
  interface
    integer(c_int) function qgcm_host_register_array(host, bytes) bind(C, name="qgcm_host_register")
      import
      real(c_double), intent(in) :: host(*)
      integer(c_int64_t), value :: bytes
    end function qgcm_host_register_array
  end interface

  ! host readers of the main program and the arrays each one reads (SURVEY.md 8b):
  integer, parameter :: GPU_FOR_VALIDS  = 1      ! src/valsubs.F:59-67
  integer, parameter :: GPU_FOR_MONNC   = 2      ! src/monitor_diag.F:114-121
  integer, parameter :: GPU_FOR_OCNC    = 3      ! src/nc_subs.F:845-848
  integer, parameter :: GPU_FOR_ATNC    = 4      ! src/nc_subs.F:1077-1100
  integer, parameter :: GPU_FOR_TAVOCN  = 5      ! src/timavge.F:431-436
  integer, parameter :: GPU_FOR_TAVATM  = 6      ! src/timavge.F:284-289
  integer, parameter :: GPU_FOR_QOCDIAG = 7      ! src/qocdiag.F:393-400
  integer, parameter :: GPU_FOR_RESTART = 8      ! src/nc_subs.F:1331-1360
  integer, parameter :: GPU_FOR_PRSAMP  = 9      ! src/q-gcm.F:1939-1952, :2127-2137
  integer, parameter :: GPU_FOR_AREAVG  = 10     ! src/areasubs_diag.F:50-80
  integer, parameter :: GPU_FOR_COVOCN  = 11     ! src/covaria_diag.F:151-215
  integer, parameter :: GPU_FOR_COVATM  = 12

contains

  ! reference convention is print + stop (src/nc_subs.F:84-112, src/ocisubs.F:361-365)
  subroutine gpu_check(ierr, where)
    integer(c_int), intent(in) :: ierr
    character(len=*), intent(in) :: where
    character(kind=c_char), pointer :: msg(:)
    integer :: n
    if (ierr == 0) return
    call c_f_pointer(qgcm_last_error(), msg, [1024])
    n = 0
    do while (n < 1024)
      if (msg(n+1) == c_null_char) exit
      n = n + 1
    end do
    print *, ' libqgcm_b200 error in ', where, ': ', msg(1:n)
    stop 1
  end subroutine gpu_check

  ! ---- one field by the reference's variable name -----------------------------------------
  subroutine put(name, arr, n)
    character(len=*), intent(in) :: name
    real(c_double), intent(in) :: arr(*)
    integer, intent(in) :: n
    call gpu_check(qgcm_set_field(gpu, trim(name)//c_null_char, arr, int(n, c_int64_t)), 'set '//name)
  end subroutine put

  subroutine get(name, arr, n)
    character(len=*), intent(in) :: name
    real(c_double), intent(inout) :: arr(*)
    integer, intent(in) :: n
    call gpu_check(qgcm_get_field(gpu, trim(name)//c_null_char, arr, int(n, c_int64_t)), 'get '//name)
  end subroutine get

  ! ---- qgcm_config from the reference's modules (after radiat, eigmod: src/q-gcm.F:560-700) ----
  subroutine gpu_fill_config(cfg, device, nranks, rank)
    use parameters
    use occonst
    use atconst
    use intrfac
    use radiate
    use timinfo, only : nstr
    type(qgcm_config), intent(out) :: cfg
    integer, intent(in) :: device, nranks, rank
    integer :: k, m
    cfg%abi_version = QGCM_ABI_VERSION
    cfg%struct_bytes = int(c_sizeof(cfg), c_int32_t)
    cfg%device = device
    cfg%nranks = nranks
    cfg%rank = rank
    cfg%reserved_i = 0
    cfg%reserved_d = 0.0d0
    cfg%flags = 0
#ifdef ocean_only
    cfg%flags = ior(cfg%flags, QGCM_OCEAN_ONLY)
#endif
#ifdef atmos_only
    cfg%flags = ior(cfg%flags, QGCM_ATMOS_ONLY)
#endif
#ifdef cyclic_ocean
    cfg%flags = ior(cfg%flags, QGCM_CYCLIC_OCEAN)
#endif
#ifdef sb_hflux
    cfg%flags = ior(cfg%flags, QGCM_SB_HFLUX)
#endif
#ifdef nb_hflux
    cfg%flags = ior(cfg%flags, QGCM_NB_HFLUX)
#endif
#ifdef tau_udiff
    cfg%flags = ior(cfg%flags, QGCM_TAU_UDIFF)
#endif
#ifdef ocnc_avg_k247
    cfg%flags = ior(cfg%flags, QGCM_OCNC_AVG_K247)
#endif
    cfg%nxto = nxto;  cfg%nyto = nyto;  cfg%nlo = nlo
    cfg%nxta = nxta;  cfg%nyta = nyta;  cfg%nla = nla
    cfg%ndxr = ndxr;  cfg%nx1 = nx1;    cfg%ny1 = ny1;   cfg%nstr = nstr
    cfg%fnot = fnot;  cfg%beta = beta;  cfg%dxo = dxo;   cfg%dta = dta
    cfg%delek = delek; cfg%cdat = cdat; cfg%rhoat = rhoat; cfg%rhooc = rhooc; cfg%cpat = cpat; cfg%cpoc = cpoc
    cfg%bccoat = bccoat; cfg%bccooc = bccooc; cfg%xcexp = xcexp; cfg%ycexp = ycexp
    cfg%xlamda = xlamda; cfg%hmoc = hmoc; cfg%st2d = st2d; cfg%st4d = st4d
    cfg%hmat = hmat; cfg%hmamin = hmamin; cfg%ahmd = ahmd; cfg%at2d = at2d; cfg%at4d = at4d; cfg%hmadmp = hmadmp
    cfg%tsbdy = tsbdy; cfg%tnbdy = tnbdy; cfg%fspco = fspco
    cfg%Bmup = Bmup; cfg%B1down = B1down; cfg%Cmup = Cmup; cfg%C1down = C1down
    cfg%D0up = D0up; cfg%Dmup = Dmup; cfg%Dmdown = Dmdown
    cfg%bface = bface; cfg%cface = cface; cfg%dface = dface
    ! matrices travel in Fortran element order with leading dimension nl (not QGCM_NLMAX)
    cfg%Aup = 0.0d0; cfg%Adown = 0.0d0
    do m = 1, nla-1
      do k = 1, nla
        cfg%Aup(k + nla*(m-1)) = Aup(k,m)
        cfg%Adown(k + nla*(m-1)) = Adown(k,m)
      end do
    end do
    cfg%Bup = 0.0d0; cfg%Cup = 0.0d0; cfg%Dup = 0.0d0; cfg%rbetat = 0.0d0; cfg%aface = 0.0d0
    cfg%Bup(1:nla) = Bup; cfg%Cup(1:nla) = Cup; cfg%Dup(1:nla) = Dup
    cfg%rbetat(1:nla-1) = rbetat; cfg%aface(1:nla-1) = aface
    cfg%hoc = 0.0d0; cfg%gpoc = 0.0d0; cfg%ah2oc = 0.0d0; cfg%ah4oc = 0.0d0; cfg%toc = 0.0d0
    cfg%hoc(1:nlo) = hoc; cfg%gpoc(1:nlo-1) = gpoc; cfg%ah2oc(1:nlo) = ah2oc; cfg%ah4oc(1:nlo) = ah4oc; cfg%toc(1:nlo) = toc
    cfg%hat = 0.0d0; cfg%gpat = 0.0d0; cfg%ah4at = 0.0d0; cfg%tat = 0.0d0
    cfg%hat(1:nla) = hat; cfg%gpat(1:nla-1) = gpat; cfg%ah4at(1:nla) = ah4at; cfg%tat(1:nla) = tat
    cfg%amatoc = 0.0d0; cfg%ctl2moc = 0.0d0; cfg%ctm2loc = 0.0d0; cfg%rdm2oc = 0.0d0
    do m = 1, nlo
      cfg%rdm2oc(m) = rdm2oc(m)
      do k = 1, nlo
        cfg%amatoc(k + nlo*(m-1)) = amatoc(k,m)
        cfg%ctl2moc(k + nlo*(m-1)) = ctl2moc(k,m)
        cfg%ctm2loc(k + nlo*(m-1)) = ctm2loc(k,m)
      end do
    end do
    cfg%amatat = 0.0d0; cfg%ctl2mat = 0.0d0; cfg%ctm2lat = 0.0d0; cfg%rdm2at = 0.0d0
    do m = 1, nla
      cfg%rdm2at(m) = rdm2at(m)
      do k = 1, nla
        cfg%amatat(k + nla*(m-1)) = amatat(k,m)
        cfg%ctl2mat(k + nla*(m-1)) = ctl2mat(k,m)
        cfg%ctm2lat(k + nla*(m-1)) = ctm2lat(k,m)
      end do
    end do
  end subroutine gpu_fill_config

  ! ---- start-up: call once after homsol (src/q-gcm.F:976) -----------------------------------
  ! Creates the device model, page-locks the state arrays (restart and output then run at PCIe
  ! speed), uploads constants and initial state and lets the device recompute everything that
  ! depends on them, so host and device constants agree by construction.
  subroutine gpu_init(device)
    use parameters
    use ocstate
    use atstate
    use intrfac
    integer, intent(in) :: device
    type(qgcm_config) :: cfg
    call gpu_fill_config(cfg, device, 1, 0)
    call gpu_check(qgcm_create(cfg, gpu), 'qgcm_create')
#ifndef atmos_only
    call gpu_check(qgcm_host_register_array(po,   int(8*size(po),   c_int64_t)), 'register po')
    call gpu_check(qgcm_host_register_array(pom,  int(8*size(pom),  c_int64_t)), 'register pom')
    call gpu_check(qgcm_host_register_array(qo,   int(8*size(qo),   c_int64_t)), 'register qo')
    call gpu_check(qgcm_host_register_array(qom,  int(8*size(qom),  c_int64_t)), 'register qom')
    call gpu_check(qgcm_host_register_array(sst,  int(8*size(sst),  c_int64_t)), 'register sst')
    call gpu_check(qgcm_host_register_array(sstm, int(8*size(sstm), c_int64_t)), 'register sstm')
#endif
#ifndef ocean_only
    call gpu_check(qgcm_host_register_array(pa,   int(8*size(pa),   c_int64_t)), 'register pa')
    call gpu_check(qgcm_host_register_array(pam,  int(8*size(pam),  c_int64_t)), 'register pam')
    call gpu_check(qgcm_host_register_array(qa,   int(8*size(qa),   c_int64_t)), 'register qa')
    call gpu_check(qgcm_host_register_array(qam,  int(8*size(qam),  c_int64_t)), 'register qam')
#endif
    call gpu_upload_state()
    call gpu_check(qgcm_constr(gpu), 'constr')                 ! src/q-gcm.F:711
#ifndef atmos_only
    call gpu_check(qgcm_qcomp_ocean(gpu), 'qcomp_ocean')       ! src/q-gcm.F:719-732
#endif
#ifndef ocean_only
    call gpu_check(qgcm_qcomp_atmos(gpu), 'qcomp_atmos')       ! src/q-gcm.F:734-746
#endif
    call gpu_check(qgcm_xforc(gpu), 'xforc')                   ! src/q-gcm.F:870
    call gpu_check(qgcm_homsol(gpu), 'homsol')                 ! src/q-gcm.F:976
    call gpu_check(qgcm_tavini(gpu), 'tavini')                 ! src/q-gcm.F:1190
  end subroutine gpu_init

  subroutine gpu_finalize()
    if (c_associated(gpu)) call gpu_check(qgcm_destroy(gpu), 'qgcm_destroy')
    gpu = c_null_ptr
  end subroutine gpu_finalize

  ! ---- every input field, by the reference's variable name ---------------------------------
  subroutine gpu_upload_state()
    use parameters
    use ocstate
    use atstate
    use intrfac
    use occonst, only : ddynoc
    use atconst, only : ddynat, dtopat, xc1ast
#ifndef atmos_only
    call put('po',     po,     size(po))
    call put('pom',    pom,    size(pom))
    call put('sst',    sst,    size(sst))
    call put('sstm',   sstm,   size(sstm))
    call put('tauxo',  tauxo,  size(tauxo))
    call put('tauyo',  tauyo,  size(tauyo))
    call put('fnetoc', fnetoc, size(fnetoc))
    call put('ddynoc', ddynoc, size(ddynoc))
    call put('sstbar', sstbar, size(sstbar))
    call put('entoc',  entoc,  size(entoc))
#endif
#ifndef ocean_only
    call put('pa',     pa,     size(pa))
    call put('pam',    pam,    size(pam))
    call put('ast',    ast,    size(ast))
    call put('astm',   astm,   size(astm))
    call put('hmixa',  hmixa,  size(hmixa))
    call put('hmixam', hmixam, size(hmixam))
    call put('ddynat', ddynat, size(ddynat))
    call put('dtopat', dtopat, size(dtopat))
    call put('xc1ast', xc1ast, size(xc1ast))
    call put('astbar', astbar, size(astbar))
    call put('entat',  entat,  size(entat))
#endif
  end subroutine gpu_upload_state

  ! ---- refresh exactly the host arrays a host reader is about to read ----------------------
  subroutine gpu_download(reader)
    use parameters
    use ocstate
    use atstate
    use intrfac
    integer, intent(in) :: reader
    select case (reader)
    case (GPU_FOR_VALIDS)                  ! or qgcm_valids on the device (INTEGRATION.md 3a)
#ifndef atmos_only
      call get('po', po, size(po));  call get('qo', qo, size(qo))
      call get('wekto', wekto, size(wekto));  call get('sst', sst, size(sst))
#endif
#ifndef ocean_only
      call get('pa', pa, size(pa));  call get('qa', qa, size(qa));  call get('wekta', wekta, size(wekta))
      call get('ast', ast, size(ast));  call get('tauxa', tauxa, size(tauxa));  call get('tauya', tauya, size(tauya))
#endif
    case (GPU_FOR_MONNC)                   ! or qgcm_monnc_ocean / qgcm_monnc_atmos on the device
#ifndef atmos_only
      call get('po', po, size(po));  call get('pom', pom, size(pom));  call get('qo', qo, size(qo))
      call get('wekto', wekto, size(wekto));  call get('wekpo', wekpo, size(wekpo));  call get('entoc', entoc, size(entoc))
      call get('sst', sst, size(sst));  call get('tauxo', tauxo, size(tauxo));  call get('tauyo', tauyo, size(tauyo))
#endif
#ifndef ocean_only
      call get('pa', pa, size(pa));  call get('pam', pam, size(pam));  call get('qa', qa, size(qa))
      call get('wekta', wekta, size(wekta));  call get('wekpa', wekpa, size(wekpa));  call get('entat', entat, size(entat))
      call get('ast', ast, size(ast));  call get('hmixa', hmixa, size(hmixa))
      call get('tauxa', tauxa, size(tauxa));  call get('tauya', tauya, size(tauya))
      call get('uekat', uekat, size(uekat));  call get('vekat', vekat, size(vekat))
#endif
      call gpu_sync_scalars()
    case (GPU_FOR_OCNC)                    ! or qgcm_get_field_sub with nsko (1/nsko**2 of the bytes)
      call get('po', po, size(po));  call get('qo', qo, size(qo));  call get('sst', sst, size(sst))
      call get('wekto', wekto, size(wekto));  call get('tauxo', tauxo, size(tauxo));  call get('tauyo', tauyo, size(tauyo))
    case (GPU_FOR_ATNC)
      call get('pa', pa, size(pa));  call get('qa', qa, size(qa));  call get('ast', ast, size(ast))
      call get('tauxa', tauxa, size(tauxa));  call get('tauya', tauya, size(tauya))
      call get('hmixa', hmixa, size(hmixa));  call get('wekta', wekta, size(wekta))
    case (GPU_FOR_TAVOCN)                  ! or qgcm_tavocn: the sums stay on the device
      call get('po', po, size(po));  call get('sst', sst, size(sst));  call get('wekto', wekto, size(wekto))
      call get('tauxo', tauxo, size(tauxo));  call get('tauyo', tauyo, size(tauyo));  call get('fnetoc', fnetoc, size(fnetoc))
    case (GPU_FOR_TAVATM)
      call get('pa', pa, size(pa));  call get('ast', ast, size(ast));  call get('wekta', wekta, size(wekta))
      call get('tauxa', tauxa, size(tauxa));  call get('tauya', tauya, size(tauya));  call get('fnetat', fnetat, size(fnetat))
    case (GPU_FOR_QOCDIAG)                 ! between oml and qgostep (src/q-gcm.F:1237-1239); or qgcm_qocdiag
      call get('po', po, size(po));  call get('pom', pom, size(pom));  call get('qo', qo, size(qo));  call get('qom', qom, size(qom))
      call get('wekpo', wekpo, size(wekpo));  call get('entoc', entoc, size(entoc))
    case (GPU_FOR_RESTART)                 ! page-locked arrays: one batched DMA (see gpu_restart_download)
      call gpu_restart_download()
    case (GPU_FOR_PRSAMP)
#ifndef atmos_only
      call get('po', po, size(po));  call get('qo', qo, size(qo));  call get('sst', sst, size(sst))
#endif
#ifndef ocean_only
      call get('pa', pa, size(pa));  call get('qa', qa, size(qa));  call get('ast', ast, size(ast))
      call get('hmixa', hmixa, size(hmixa));  call get('uekat', uekat, size(uekat));  call get('vekat', vekat, size(vekat))
#endif
    case (GPU_FOR_AREAVG)
      call get('sst', sst, size(sst))
#ifndef ocean_only
      call get('ast', ast, size(ast))
#endif
    case (GPU_FOR_COVOCN)
      call get('po', po, size(po));  call get('sst', sst, size(sst))
    case (GPU_FOR_COVATM)
      call get('pa', pa, size(pa));  call get('ast', ast, size(ast))
    end select
  end subroutine gpu_download

  ! the ten restart fields of resave_nc (src/nc_subs.F:1331-1360).  The arrays were page-locked in
  ! gpu_init, so each qgcm_get_field is a DMA at PCIe speed (measured 52 GB/s on the B200 box;
  ! 14 GB/s without the registration).  qgcm_get_fields would batch them under one
  ! synchronisation, but it takes C addresses, and c_loc needs the TARGET attribute the
  ! reference's module arrays do not have -- ten synchronisations cost microseconds.
  subroutine gpu_restart_download()
    use parameters
    use ocstate
    use atstate
    use intrfac
#ifndef atmos_only
    call get('po', po, size(po));  call get('pom', pom, size(pom))
    call get('sst', sst, size(sst));  call get('sstm', sstm, size(sstm))
#endif
#ifndef ocean_only
    call get('pa', pa, size(pa));  call get('pam', pam, size(pam))
    call get('ast', ast, size(ast));  call get('astm', astm, size(astm))
    call get('hmixa', hmixa, size(hmixa));  call get('hmixam', hmixam, size(hmixam))
#endif
  end subroutine gpu_restart_download

  ! ---- scalar state the step mutates: device -> the reference's module variables -----------
  ! (src/ochomog_data.F:57-69, src/athomog_data.F:47-55, src/monitor_data.F:41-61)
  subroutine gpu_sync_scalars()
    use parameters
    use ochomog
    use athomog
    use monitor
    type(qgcm_scalars) :: s
    call gpu_check(qgcm_get_scalars(gpu, s), 'get_scalars')
#ifndef atmos_only
    xon = s%xon(1:nlo-1);  dpioc = s%dpioc(1:nlo-1);  dpiocp = s%dpiocp(1:nlo-1)
    ermaso = s%ermaso(1:nlo-1);  emfroc = s%emfroc(1:nlo-1)
    cfraoc = s%cfraoc;  centoc = s%centoc
    ttmads = s%ttmads;  vfmads = s%vfmads;  ttmdfs = s%ttmdfs
    ttmadn = s%ttmadn;  vfmadn = s%vfmadn;  ttmdfn = s%ttmdfn
#  ifdef cyclic_ocean
    ocncs = s%ocncs(1:nlo);  ocncn = s%ocncn(1:nlo);  ocncsp = s%ocncsp(1:nlo);  ocncnp = s%ocncnp(1:nlo)
    enisoc = s%enisoc(1:nlo-1);  eninoc = s%eninoc(1:nlo-1)
    ajisoc = s%ajisoc(1:nlo);  ajinoc = s%ajinoc(1:nlo)
    ap3soc = s%ap3soc(1:nlo);  ap3noc = s%ap3noc(1:nlo);  ap5soc = s%ap5soc(1:nlo);  ap5noc = s%ap5noc(1:nlo)
    txisoc = s%txisoc;  txinoc = s%txinoc;  bdrins = s%bdrins;  bdrinn = s%bdrinn
#  endif
#endif
#ifndef ocean_only
    xan = s%xan(1:nla-1);  dpiat = s%dpiat(1:nla-1);  dpiatp = s%dpiatp(1:nla-1)
    atmcs = s%atmcs(1:nla);  atmcn = s%atmcn(1:nla);  atmcsp = s%atmcsp(1:nla);  atmcnp = s%atmcnp(1:nla)
    enisat = s%enisat(1:nla-1);  eninat = s%eninat(1:nla-1)
    ajisat = s%ajisat(1:nla);  ajinat = s%ajinat(1:nla);  ap5sat = s%ap5sat(1:nla);  ap5nat = s%ap5nat(1:nla)
    txisat = s%txisat;  txinat = s%txinat
    ermasa = s%ermasa(1:nla-1);  emfrat = s%emfrat(1:nla-1)
    cfraat = s%cfraat;  centat = s%centat
    arlaav = s%arlaav;  slhfav = s%slhfav;  oradav = s%oradav;  arocav = s%arocav
#endif
  end subroutine gpu_sync_scalars

end module qgcm_cuda
