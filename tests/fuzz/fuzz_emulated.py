"""Fuzz the SHIPPED kernels on the CPU: random geometries through the real launchers under tests/host_shim/cuda_emul.h, checked against the
oracle (FIR bank incl. forced tilings and the generic path, fused DDC bank with chunk offsets, shift_addition / shift_math banks, fractional
decimator).  usage: python tests/fuzz/fuzz_emulated.py [seed] [seconds]   -- test infrastructure, needs g++ and the CUDA headers only."""
import sys, time, ctypes as C, numpy as np
from pathlib import Path as _P
_ROOT = str(_P(__file__).resolve().parents[2])
sys.path.insert(0, _ROOT); sys.path.insert(0, _ROOT + '/tests/host_shim')
import emul_build as eb
from pathlib import Path
from oracle.pyoracle import Oracle, rel_rms
o=Oracle()
import tempfile
out=Path(tempfile.mkdtemp(prefix='fuzz_emul_'))
fir,_=eb.build_file(out,'fir_decimate.cu'); ddc,_=eb.build_file(out,'ddc_bank.cu'); sh,_=eb.build_file(out,'shift.cu'); au,_=eb.build_file(out,'audio.cu',host_c=("csdr_b200/host/firdes.c",))
P=lambda a:a.ctypes.data
def aligned(shape,dtype):
    n=int(np.prod(shape)); it=np.dtype(dtype).itemsize; raw=np.zeros(n*it+32,np.uint8); off=(-raw.ctypes.data)%16
    return raw[off:off+n*it].view(dtype).reshape(shape)
def cplx(rng,*s): return (rng.uniform(-1,1,s)+1j*rng.uniform(-1,1,s)).astype(np.complex64)
rng=np.random.default_rng(int(sys.argv[1]) if len(sys.argv)>1 else 0)
t_end=time.time()+float(sys.argv[2]) if len(sys.argv)>2 else time.time()+120
it=0; worst={}
while time.time()<t_end:
    it+=1
    kind=rng.integers(0,4)
    if kind==0:   # FIR bank
        D,T=[(10,199),(10,79),(50,801),(10,int(rng.integers(1,200))),(50,int(rng.integers(1,900))),(int(rng.integers(1,40)),int(rng.integers(1,300)))][rng.integers(0,6)]
        n=int(rng.integers(1,40000)); ch=int(rng.integers(1,4)); variant=int(rng.integers(-1,8)) if (D==10 and 80<T<=200) else -1
        stride=n+(n&1); x=aligned((ch,stride),np.complex64); x[:,:n]=cplx(rng,ch,n); x[:,n:]=np.nan
        taps=rng.uniform(-1,1,T).astype(np.float32); n_out=(n-T)//D+1 if n>=T else 0; ostride=max(n_out+(n_out&1),2)
        y=aligned((ch,ostride),np.complex64); y[:]=np.nan; fp=taps.ctypes.data_as(C.c_void_p)
        rc=fir.emul_launch_fir_decimate_bank(P(x),stride,P(y),ostride,ch,n,D,fp,fp,0,T,variant)
        assert rc==n_out,(D,T,n,variant,rc,fir.emul_last_error())
        for c in range(ch):
            w=o.fir_decimate_cc(np.ascontiguousarray(x[c,:n]),D,taps)
            if n_out:
                e=np.abs(y[c,:n_out]-w).max()/max(np.abs(taps).sum(),1e-9); worst['fir']=max(worst.get('fir',0),e)
                assert e<3e-6,("fir",D,T,n,variant,c,e)
    elif kind==1:  # fused ddc bank with offset/chunk/streaming
        D,bw=[(50,0.005),(10,0.0201),(10,0.05)][rng.integers(0,3)]; T=o.firdes_filter_len(bw); taps=o.firdes_lowpass_f(T,0.5/D)
        n=int(rng.integers(T,30000))&~1; chn=int(rng.integers(1,40)); chunk=[1024,1000,4096,256][rng.integers(0,4)]; offset=int(rng.integers(0,chunk)); demod=int(rng.integers(0,2))
        if (offset&1): offset-=1
        rates=rng.uniform(-0.5,0.5,chn).astype(np.float32)
        # absolute stream: chunk boundaries at multiples of chunk; block starts `offset` into a chunk -> oracle: prepend offset dummy samples processed with the same phase
        x=aligned(n,np.complex64); x[:]=cplx(rng,n)*0.5
        params=np.array([o.shift_addition_init(float(r)) for r in rates],np.float32); ph0=rng.uniform(-3,3,chn).astype(np.float32); ph=ph0.copy()
        n_out=(n-T)//D+1; stride=n_out+(n_out&1)
        res={}
        for dm in ((0,1) if demod else (0,)):
            ph=ph0.copy(); y=np.zeros((chn,stride),np.float32 if dm else np.complex64); lo=np.zeros(chn,np.complex64); la=C.c_int(0)
            sb=ddc.emul_ddc_bank_scratch_bytes(chn,n,chunk,offset); scr=np.zeros(sb+64,np.uint8)
            rc=ddc.emul_launch_ddc_bank(P(x),n,chn,P(params),P(ph),chunk,offset,D,taps.ctypes.data_as(C.c_void_p),T,dm,P(y),stride,None,P(lo) if dm else None,P(scr),sb,C.addressof(la))
            assert rc==n_out,(rc,ddc.emul_last_error()); res[dm]=y
        for c in range(min(chn,3)):
            r=float(rates[c])
            # oracle: stream = [offset zeros | x], cut into calls from the chunk start, starting phase ph0
            full=np.concatenate([np.zeros(offset,np.complex64),x]); s_,_=o.shift_addition_cc(full,r,float(ph0[c]),chunk); s_=s_[offset:]
            base=o.fir_decimate_cc(s_,D,taps)
            e=rel_rms(res[0][c,:n_out],base); worst['ddc']=max(worst.get('ddc',0),e); assert e<3e-6,("ddc",D,n,chunk,offset,c,e)
            if demod:   # the fused discriminator == the oracle's discriminator on the kernel's own baseband, bit for bit (noise input amplifies any baseband ulp)
                assert np.array_equal(res[1][c,:n_out],o.fmdemod_quadri_cf(np.ascontiguousarray(res[0][c,:n_out]))[0]),("ddc demod",D,n,chunk,offset,c)
    elif kind==2:  # shift banks
        n=int(rng.integers(1,30000)); chunk=int([0,1,37,1000,1024,4096][rng.integers(0,6)]); chn=int(rng.integers(1,5))
        rates=rng.uniform(-0.5,0.5,chn).astype(np.float32); x=cplx(rng,n); ph0=rng.uniform(-30,30,chn).astype(np.float32); ph=ph0.copy()
        params=np.array([o.shift_addition_init(float(r)) for r in rates],np.float32); y=np.zeros((chn,n),np.complex64)
        sb=sh.emul_shift_bank_scratch_bytes(chn,n,chunk); scr=np.zeros(sb+16,np.uint8)
        assert sh.emul_launch_shift_addition_bank(P(x),0,P(y),n,chn,n,P(params),P(ph),chunk,P(scr),sb)>=0
        for c in range(chn):
            w,wp=o.shift_addition_cc(x,float(rates[c]),float(ph0[c]),chunk or None)
            assert np.float32(wp)==ph[c],("shift phase",n,chunk,c,wp,ph[c]); e=rel_rms(y[c],w); worst['shift']=max(worst.get('shift',0),e); assert e<2e-7,("shift",n,chunk,e)
        ph=ph0.copy(); sbm=sh.emul_shift_math_scratch_bytes(chn,n); scr=np.zeros(sbm+16,np.uint8)
        assert sh.emul_launch_shift_math_bank(P(x),0,P(y),n,chn,n,P(rates),P(ph),P(scr),sbm)>=0
        for c in range(chn):
            w,wp=o.shift_math_cc(x,float(rates[c]),float(ph0[c])); assert np.float32(wp).view(np.uint32)==ph[c].view(np.uint32),("math phase",n,c); assert rel_rms(y[c],w)<2e-7
    else:  # fracdec + fastagc
        n=int(rng.integers(300,30000)); rate=float(np.float32(rng.uniform(1.01,20))); pts=int([2,4,8,12,16,32][rng.integers(0,6)]); chn=int(rng.integers(1,4))
        if n<=2*pts+4: continue
        x=rng.uniform(-1,1,(chn,n)).astype(np.float32); cap=int(n/rate)+8; y=np.zeros((chn,cap),np.float32)
        st=np.zeros((chn,3),np.int32); st[:,0]=np.array([pts//2-1],np.float32).view(np.int32)[0]
        sb=au.emul_fracdec_scratch_bytes(chn,n,rate); scr=np.zeros(sb+16,np.uint8)
        assert au.emul_launch_fractional_decimator_bank(P(x),n,P(y),cap,chn,n,rate,pts,None,0,P(st),P(scr),sb)>=0,au.emul_last_error()
        for c in range(chn):
            w=o.fractional_decimator_ff(x[c],rate,pts); assert st[c,2]==w.size,("fracdec count",n,rate,pts,st[c,2],w.size); assert np.array_equal(y[c,:w.size],w),("fracdec",n,rate,pts)
print("iterations",it,"worst",worst)
