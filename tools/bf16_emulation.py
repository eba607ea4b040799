import os, sys; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import sr_wavenet_b200.synth as synth
from oracle import srwn_oracle as orc

def bf16(a):
    a = np.asarray(a, np.float32)
    u = a.view(np.uint32).astype(np.uint64)
    r = ((u + 0x7FFF + ((u >> 16) & 1)) >> 16) << 16
    return r.astype(np.uint32).view(np.float32).astype(np.float64)

def emul(w, x, enc, dil, P, tanh_err=0.0, rng=None):
    pre='WaveNetAutoEncoder/Decoder/'
    W=lambda n: w[pre+n].astype(np.float64)
    B,T=x.shape
    xs = np.concatenate([np.zeros((B,1)), x[:,:-1]],1)[:,:,None]
    h = orc.dilated_causal_conv1d_layer(xs, W('causal_conv_Kernel'), W('causal_conv_Bias'),1)
    total=0
    n=len(dil)
    for i,d in enumerate(dil):
        cn='conv1d' if i==0 else 'conv1d_%d'%(3*i)
        cond = enc@W(cn+'/kernel')[0]+W(cn+'/bias')
        h = h + np.repeat(cond,P,axis=1)
        name='dilated_conv_%d'%i
        fk=bf16(W('%s_filter/%s_Kernel'%(name,name))); fb=W('%s_filter/%s_Bias'%(name,name))
        hb=bf16(h)
        a = orc.dilated_causal_conv1d(hb, fk, d)+fb.reshape(1,1,-1)
        f=np.tanh(a)
        if tanh_err: f = f*(1+rng.uniform(-tanh_err,tanh_err,f.shape))
        g = 0.5+0.5*np.tanh(0.5*f)
        if tanh_err: g = 0.5+0.5*np.tanh(0.5*f)*(1+rng.uniform(-tanh_err,tanh_err,f.shape))
        c=bf16(f*g)
        res=c@bf16(W('conv1d_%d/kernel'%(3*i+1))[0])+W('conv1d_%d/bias'%(3*i+1))
        total = total + c@bf16(W('conv1d_%d/kernel'%(3*i+2))[0])+W('conv1d_%d/bias'%(3*i+2))
        h=(h+res)*orc.SQRT_HALF
    total=bf16(np.maximum(total,0))
    hid=bf16(np.maximum(total@bf16(W('conv1d_%d/kernel'%(3*n))[0])+W('conv1d_%d/bias'%(3*n)),0))
    return hid@bf16(W('conv1d_%d/kernel'%(3*n+1))[0])+W('conv1d_%d/bias'%(3*n+1))

dil=synth.DEFAULT_DILATIONS
tw=synth.make_teacher_weights(dil)
B,T,P=2,4096,128
x=synth.synthetic_audio(B,T).astype(np.float64); enc=synth.synthetic_encoding(B,T//P).astype(np.float64)
ref=orc.teacher_decoder_logits({k:v.astype(np.float64) for k,v in tw.items()},x,enc,dil,P)
rng=np.random.default_rng(0)
for te in (0.0, 2**-11, 2**-9):
    e=emul(tw,x,enc,dil,P,te,rng)
    d=np.abs(e-ref)
    print("tanh_err %.1e: max|dlogits| %.4f mean %.5f  logits absmax %.2f"%(te,d.max(),d.mean(),np.abs(ref).max()))
nll_ref=orc.discretized_mix_logistic_loss(x[:,:,None],ref,False); nll=orc.discretized_mix_logistic_loss(x[:,:,None],e,False)
print("nll max diff", np.abs(nll-nll_ref).max(), "sum rel", abs(nll.sum()-nll_ref.sum())/abs(nll_ref.sum()))

def emul2(w, x, enc, dil, P, q_h=True, q_c=True, q_w=True, q_head=True, q_hw=True, f16head=False):
    pre='WaveNetAutoEncoder/Decoder/'
    W=lambda n: w[pre+n].astype(np.float64)
    Q=lambda a,on: bf16(a) if on else a
    B,T=x.shape
    xs = np.concatenate([np.zeros((B,1)), x[:,:-1]],1)[:,:,None]
    h = orc.dilated_causal_conv1d_layer(xs, W('causal_conv_Kernel'), W('causal_conv_Bias'),1)
    total=0; n=len(dil)
    for i,d in enumerate(dil):
        cn='conv1d' if i==0 else 'conv1d_%d'%(3*i)
        h = h + np.repeat(enc@W(cn+'/kernel')[0]+W(cn+'/bias'),P,axis=1)
        name='dilated_conv_%d'%i
        fk=Q(W('%s_filter/%s_Kernel'%(name,name)),q_w); fb=W('%s_filter/%s_Bias'%(name,name))
        f=np.tanh(orc.dilated_causal_conv1d(Q(h,q_h), fk, d)+fb.reshape(1,1,-1))
        c=Q(f*orc.sigmoid(f),q_c)
        res=c@Q(W('conv1d_%d/kernel'%(3*i+1))[0],q_w)+W('conv1d_%d/bias'%(3*i+1))
        total = total + c@Q(W('conv1d_%d/kernel'%(3*i+2))[0],q_w)+W('conv1d_%d/bias'%(3*i+2))
        h=(h+res)*orc.SQRT_HALF
    total=Q(np.maximum(total,0),q_head)
    hid=Q(np.maximum(total@Q(W('conv1d_%d/kernel'%(3*n))[0],q_hw)+W('conv1d_%d/bias'%(3*n)),0),q_head)
    return hid@Q(W('conv1d_%d/kernel'%(3*n+1))[0],q_hw)+W('conv1d_%d/bias'%(3*n+1))
for name,kw in [("all",{}),("no h",dict(q_h=False)),("no c",dict(q_c=False)),("no w",dict(q_w=False)),("no head act",dict(q_head=False)),("no head w",dict(q_hw=False)),("only h",dict(q_c=False,q_w=False,q_head=False,q_hw=False)),("only c",dict(q_h=False,q_w=False,q_head=False,q_hw=False)),("only w",dict(q_h=False,q_c=False,q_head=False,q_hw=False)),("only head",dict(q_h=False,q_c=False,q_w=False))]:
    e=emul2(tw,x,enc,dil,P,**kw); d=np.abs(e-ref); print("%-12s max %.4f mean %.5f"%(name,d.max(),d.mean()))
