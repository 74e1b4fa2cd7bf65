// walk_list_sim.cpp -- CPU study for the list walk (design tool, not on the product path).
//
// Question: how many of the (cell, target) visits of the reference's per-particle acceptance rule
// (F/gravOctree_Single.jl:265) can be decided for a whole warp of 32 key-adjacent targets at once from the warp's
// bounding box - "every lane accepts" (the cell goes to a monopole list that all lanes evaluate in a tight loop) or
// "every lane opens" (the children are tested next, again for the whole warp) - and how much is left for the masked
// per-lane walk?  Decisions stay per particle: the group tests are conservative proofs, anything unproven is "mixed".
//
// build / run:  g++ -O2 -o /tmp/walk_list_sim tools/walk_list_sim.cpp && /tmp/walk_list_sim 1000000 [group 32] [T 12]
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <algorithm>
#include <random>
using namespace std;
struct Node { double c[3], L, com[3], M, rad; int child[8]; int nch, part; double lo[3], hi[3]; };
static vector<Node> nodes; static vector<double> X, Y, Z, H;
static int build(vector<int>& idx, double cx, double cy, double cz, double L) {
    int id = nodes.size(); nodes.push_back(Node());
    { Node& n = nodes[id]; n.c[0]=cx; n.c[1]=cy; n.c[2]=cz; n.L=L; n.nch=0; n.part=-1;
      n.lo[0]=cx-L; n.lo[1]=cy-L; n.lo[2]=cz-L; n.hi[0]=cx+L; n.hi[1]=cy+L; n.hi[2]=cz+L; }
    if (idx.size()==1) { Node& n=nodes[id]; n.part=idx[0]; n.com[0]=X[idx[0]]; n.com[1]=Y[idx[0]]; n.com[2]=Z[idx[0]]; n.M=1; n.rad=0; return id; }
    vector<int> sub[8];
    for (int i: idx) { int o=(X[i]-cx>0)+2*(Y[i]-cy>0)+4*(Z[i]-cz>0); sub[o].push_back(i); }
    double m=0, s[3]={0,0,0};
    for (int o=0;o<8;++o) if(!sub[o].empty()) {
        double h=L/2; int ch=build(sub[o], cx+((o&1)?h:-h), cy+((o&2)?h:-h), cz+((o&4)?h:-h), h);
        Node& n=nodes[id]; n.child[n.nch++]=ch; m+=nodes[ch].M; for(int k=0;k<3;++k) s[k]+=nodes[ch].M*nodes[ch].com[k];
    }
    Node& n=nodes[id]; n.M=m; double r2=0; for(int k=0;k<3;++k){ n.com[k]=s[k]/m; double a=max(n.com[k]-n.lo[k], n.hi[k]-n.com[k]); r2+=a*a; } n.rad=sqrt(r2); return id;
}
static const double theta=0.576;
static inline bool accept(int i, const Node& n){ double dx=X[i]-n.com[0],dy=Y[i]-n.com[1],dz=Z[i]-n.com[2]; double d2=dx*dx+dy*dy+dz*dz; double s=2*n.L;
    bool acc = s*s/d2<theta*theta; if(acc){ double e2=0; double p[3]={X[i],Y[i],Z[i]}; for(int k=0;k<3;++k){double a=max(max(n.lo[k]-p[k],0.0),p[k]-n.hi[k]); e2+=a*a;} acc = H[i]*H[i]/e2<0.25; } return acc; }
int main(int argc,char**argv){
    int N=argc>1?atoi(argv[1]):100000; int G=argc>2?atoi(argv[2]):32; int T=argc>3?atoi(argv[3]):12; int dist=argc>4?atoi(argv[4]):0;
    mt19937_64 rng(1); uniform_real_distribution<double> U(-1,1); normal_distribution<double> Nn(0,0.3);
    if(dist==0) while((int)X.size()<N){double x=U(rng),y=U(rng),z=U(rng); if(x*x+y*y+z*z<=1){X.push_back(x);Y.push_back(y);Z.push_back(z);}}
    else for(int i=0;i<N;++i){X.push_back(Nn(rng));Y.push_back(Nn(rng));Z.push_back(Nn(rng));}
    double l=0; for(int i=0;i<N;++i) l=max(l,max(fabs(X[i]),max(fabs(Y[i]),fabs(Z[i]))));
    H.assign(N, 0.5*cbrt(50.0*3/(4*M_PI)/ (N/(4*M_PI/3))));
    vector<int> all(N); for(int i=0;i<N;++i) all[i]=i;
    nodes.reserve(2*N); int root=build(all,0,0,0,l);
    vector<int> order; { vector<int> st{root}; while(!st.empty()){int n=st.back(); st.pop_back(); if(nodes[n].part>=0) order.push_back(nodes[n].part); else for(int c=nodes[n].nch-1;c>=0;--c) st.push_back(nodes[n].child[c]);} }
    vector<int> slot(N); for(int s=0;s<N;++s) slot[order[s]]=s;
    if(dist==1){ for(int s=0;s<N;++s){int i=order[s]; int a=max(0,s-25), b=min(N-1,s+25); double m=0; for(int t=a;t<=b;++t){int j=order[t]; double d=hypot(hypot(X[i]-X[j],Y[i]-Y[j]),Z[i]-Z[j]); m=max(m,d);} H[i]=0.5*m*0.6;} }
    double u_tested=0,u_pops=0,list_cells=0,list_leaves=0,u_open=0,mixed_cells=0,mixed_leaves=0,m_dense_iter=0,m_dense_lanevis=0,m_pairs=0,tot_vis=0,m_acc_lanevis=0;
    int ngroups=0; long unsound=0; double mixed_allacc=0, mixed_allopen=0;
    for (int g0=0; g0+G<=N; g0+= G*41) {
        ++ngroups;
        double blo[3]={1e300,1e300,1e300},bhi[3]={-1e300,-1e300,-1e300},hmax=0;
        for(int k=0;k<G;++k){int i=order[g0+k]; double p[3]={X[i],Y[i],Z[i]}; for(int a=0;a<3;++a){blo[a]=min(blo[a],p[a]);bhi[a]=max(bhi[a],p[a]);} hmax=max(hmax,H[i]);}
        struct E{int n; unsigned long long m;};
        vector<int> ust{root}; vector<E> mst;
        const unsigned long long full = G==64? ~0ull : ((1ull<<G)-1);
        auto masked_test=[&](int cn, unsigned long long m, bool countiter){ // every lane of m tests cell cn
            Node& n=nodes[cn]; int pc=__builtin_popcountll(m);
            if(countiter){ ++m_dense_iter; m_dense_lanevis+=pc; }
            tot_vis+=pc;
            if(n.part>=0) return; unsigned long long om=0;
            for(int k=0;k<G;++k) if((m>>k)&1){ if(!accept(order[g0+k],n)) om|=1ull<<k; else ++m_acc_lanevis; }
            if(om) mst.push_back({cn,om}); };
        while(!ust.empty()){
            int pn=ust.back(); ust.pop_back(); ++u_pops; Node& P=nodes[pn];
            for(int c=0;c<P.nch;++c){ int cn=P.child[c]; Node& n=nodes[cn]; ++u_tested;
                double dmin2=0,dmax2=0; for(int a=0;a<3;++a){ double t1=n.com[a]-bhi[a], t2=blo[a]-n.com[a]; double mn=max(max(t1,t2),0.0); double mx=max(n.com[a]-blo[a], bhi[a]-n.com[a]); dmin2+=mn*mn; dmax2+=mx*mx; }
                if(n.part>=0){ int s=slot[n.part]; double hij=(hmax+H[n.part])/2; if((s<g0||s>=g0+G) && dmin2>4*hij*hij*(1+1e-9)){ ++list_leaves; tot_vis+=G; } else { ++mixed_leaves; masked_test(cn,full,true);} continue; }
                double s2=4*n.L*n.L; double w=n.rad+2*hmax*(1+1e-9);
                if(s2<theta*theta*dmin2*(1-1e-12) && dmin2>w*w){ ++list_cells; tot_vis+=G; for(int k=0;k<G;++k) if(!accept(order[g0+k],n)) ++unsound; }
                else if(s2>theta*theta*dmax2*(1+1e-12)){ ++u_open; tot_vis+=G; ust.push_back(cn); for(int k=0;k<G;++k) if(accept(order[g0+k],n)) ++unsound; }
                else { ++mixed_cells; int na=0; for(int k=0;k<G;++k) na+=accept(order[g0+k],n); if(na==G) ++mixed_allacc; else if(na==0) ++mixed_allopen; masked_test(cn,full,true); }
            }
        }
        while(!mst.empty()){ E e=mst.back(); mst.pop_back(); Node& P=nodes[e.n]; int pc=__builtin_popcountll(e.m);
            if(pc<=T){ m_pairs+=pc*P.nch; // pair queue: replay per-lane
                for(int k=0;k<G;++k) if((e.m>>k)&1){ vector<int> st2; for(int c=0;c<P.nch;++c) st2.push_back(P.child[c]);
                    bool firstlevel=true; (void)firstlevel; int cnt=0;
                    while(!st2.empty()){int n=st2.back(); st2.pop_back(); ++cnt; ++tot_vis; if(nodes[n].part<0 && !accept(order[g0+k],nodes[n])) for(int c=0;c<nodes[n].nch;++c) st2.push_back(nodes[n].child[c]);}
                    m_pairs+=cnt-P.nch; }
                continue; }
            for(int c=0;c<P.nch;++c) masked_test(P.child[c],e.m,true);
        }
    }
    double g=ngroups;
    printf("N=%d G=%d T=%d: per group: uniform pops %.0f children tested %.0f -> list cells %.0f + list leaves %.0f, uniform opens %.0f, mixed cells %.0f, mixed leaves %.0f\n",
           N,G,T,u_pops/g,u_tested/g,list_cells/g,list_leaves/g,u_open/g,mixed_cells/g,mixed_leaves/g);
    printf("  masked walk: dense iterations %.0f (lane-visits/target %.0f, accepted %.0f), pair visits per group %.0f (per target %.0f)\n",
           m_dense_iter/g,m_dense_lanevis/g/G,m_acc_lanevis/g/G,m_pairs/g,m_pairs/g/G);
    printf("  of the mixed cells: %.0f accepted by every lane, %.0f opened by every lane (no proof from the box)\n", mixed_allacc/g, mixed_allopen/g);
    printf("  total visits per target %.1f   unsound group decisions %ld\n", tot_vis/g/G, unsound);
}
