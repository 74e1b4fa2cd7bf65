// walk_cost_sim.cpp -- CPU design study (not on the product path): distribution of the walk's cost over the warps of 32
// key-adjacent targets (shared iterations x 100 + pair visits x 12.5 instructions, the measured averages of
// walk_pairs_kernel), and its relation to the spatial extent of the warp's targets.
// N = 1e6 uniform sphere: mean 144 000 instructions per warp, p99 1.39x, max 1.59x the mean; warps that straddle a jump of
// the key order (extent > 12 h, 4 % of the warps) cost 1.23x the mean.  Consequence for several ranks: at 8 ranks a rank owns
// 977 tiles for 1 184 resident block slots, so every slot runs ONE long work item and the launch ends with the longest
// one - 1.6x the ideal share - whatever the order (measured 1.11 .. 1.50 ms per rank against 0.73 ms).
// build / run:  g++ -O2 -o /tmp/walk_cost_sim tools/walk_cost_sim.cpp && /tmp/walk_cost_sim 1000000 [stride of sampled warps: 8]
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <algorithm>
#include <random>
using namespace std;
struct Node { double c[3], L, com[3], M; int child[8]; int nch, part; double lo[3], hi[3]; int depth; };
static vector<Node> nodes; static vector<double> X, Y, Z, H;
static int build(vector<int>& idx, double cx, double cy, double cz, double L,int depth) {
    int id = nodes.size(); nodes.push_back(Node());
    { Node& n = nodes[id]; n.c[0]=cx; n.c[1]=cy; n.c[2]=cz; n.L=L; n.nch=0; n.part=-1; n.depth=depth;
      n.lo[0]=cx-L; n.lo[1]=cy-L; n.lo[2]=cz-L; n.hi[0]=cx+L; n.hi[1]=cy+L; n.hi[2]=cz+L; }
    if (idx.size()==1) { Node& n=nodes[id]; n.part=idx[0]; n.com[0]=X[idx[0]]; n.com[1]=Y[idx[0]]; n.com[2]=Z[idx[0]]; n.M=1; return id; }
    vector<int> sub[8];
    for (int i: idx) { int o=(X[i]-cx>0)+2*(Y[i]-cy>0)+4*(Z[i]-cz>0); sub[o].push_back(i); }
    double m=0, s[3]={0,0,0};
    for (int o=0;o<8;++o) if(!sub[o].empty()) {
        double h=L/2; int ch=build(sub[o], cx+((o&1)?h:-h), cy+((o&2)?h:-h), cz+((o&4)?h:-h), h, depth+1);
        Node& n=nodes[id]; n.child[n.nch++]=ch; m+=nodes[ch].M; for(int k=0;k<3;++k) s[k]+=nodes[ch].M*nodes[ch].com[k];
    }
    Node& n=nodes[id]; n.M=m; for(int k=0;k<3;++k) n.com[k]=s[k]/m; return id;
}
static const double theta=0.576;
static inline bool accept(int i, const Node& n){ double dx=X[i]-n.com[0],dy=Y[i]-n.com[1],dz=Z[i]-n.com[2]; double d2=dx*dx+dy*dy+dz*dz; double s=2*n.L;
    bool acc = s*s/d2<theta*theta; if(acc){ double e2=0; double p[3]={X[i],Y[i],Z[i]}; for(int k=0;k<3;++k){double a=max(max(n.lo[k]-p[k],0.0),p[k]-n.hi[k]); e2+=a*a;} acc = H[i]*H[i]/e2<0.25; } return acc; }
int main(int argc,char**argv){
    int N=argc>1?atoi(argv[1]):1000000; const int G=32, T=12; int stride=argc>2?atoi(argv[2]):8;
    mt19937_64 rng(1); uniform_real_distribution<double> U(-1,1);
    while((int)X.size()<N){double x=U(rng),y=U(rng),z=U(rng); if(x*x+y*y+z*z<=1){X.push_back(x);Y.push_back(y);Z.push_back(z);}}
    double l=0; for(int i=0;i<N;++i) l=max(l,max(fabs(X[i]),max(fabs(Y[i]),fabs(Z[i]))));
    double h0=0.5*cbrt(50.0*3/(4*M_PI)/ (N/(4*M_PI/3))); H.assign(N,h0);
    vector<int> all(N); for(int i=0;i<N;++i) all[i]=i;
    nodes.reserve(2*N); int root=build(all,0,0,0,l,0);
    vector<int> order; { vector<int> st{root}; while(!st.empty()){int n=st.back(); st.pop_back(); if(nodes[n].part>=0) order.push_back(nodes[n].part); else for(int c=nodes[n].nch-1;c>=0;--c) st.push_back(nodes[n].child[c]);} }
    struct R{double cost, ext; int g0;}; vector<R> res;
    for (int g0=0; g0+G<=N; g0+=G*stride) {
        double blo[3]={1e300,1e300,1e300},bhi[3]={-1e300,-1e300,-1e300};
        for(int k=0;k<G;++k){int i=order[g0+k]; double p[3]={X[i],Y[i],Z[i]}; for(int a=0;a<3;++a){blo[a]=min(blo[a],p[a]);bhi[a]=max(bhi[a],p[a]);}}
        double ext=max(bhi[0]-blo[0],max(bhi[1]-blo[1],bhi[2]-blo[2]))/h0;
        struct E{int n; unsigned m;}; vector<E> st; st.push_back({root,0xffffffffu});
        double dense=0,pairs=0;
        while(!st.empty()){ E e=st.back(); st.pop_back(); Node& P=nodes[e.n]; int pc=__builtin_popcount(e.m);
            if(pc<=T){ for(int k=0;k<G;++k) if((e.m>>k)&1){ vector<int> s2; for(int c=0;c<P.nch;++c) s2.push_back(P.child[c]);
                    while(!s2.empty()){int n=s2.back(); s2.pop_back(); ++pairs; if(nodes[n].part<0 && !accept(order[g0+k],nodes[n])) for(int c=0;c<nodes[n].nch;++c) s2.push_back(nodes[n].child[c]);} }
                continue; }
            for(int c=0;c<P.nch;++c){ Node& n=nodes[P.child[c]]; ++dense; if(n.part>=0) continue; unsigned om=0;
                for(int k=0;k<G;++k) if((e.m>>k)&1) if(!accept(order[g0+k],n)) om|=1u<<k;
                if(om) st.push_back({P.child[c],om}); } }
        res.push_back({dense*100+pairs*12.5, ext, g0});
    }
    sort(res.begin(),res.end(),[](const R&a,const R&b){return a.cost<b.cost;});
    double mean=0; for(auto&r:res) mean+=r.cost; mean/=res.size();
    auto P=[&](double q){return res[(size_t)(q*(res.size()-1))];};
    printf("warps sampled %zu: cost (instr est.) mean %.0f p50 %.0f p90 %.0f p99 %.0f p99.9 %.0f max %.0f  (max/mean %.2f)\n",res.size(),mean,P(.5).cost,P(.9).cost,P(.99).cost,P(.999).cost,res.back().cost,res.back().cost/mean);
    printf("bbox extent / h: p50-cost warp %.1f, p90 %.1f, p99 %.1f, p99.9 %.1f, max %.1f\n",P(.5).ext,P(.9).ext,P(.99).ext,P(.999).ext,res.back().ext);
    // correlation: mean cost by extent bucket
    for(double lo: {0.0,3.0,4.0,6.0,8.0,12.0,20.0}){ double s=0;int c=0; double hi = lo==0?3:(lo==3?4:(lo==4?6:(lo==6?8:(lo==8?12:(lo==12?20:1e9))))); for(auto&r:res) if(r.ext>=lo&&r.ext<hi){s+=r.cost;++c;} if(c) printf("  extent [%.0f,%.0f) h: %d warps (%.1f %%), mean cost %.0f\n",lo,hi,c,100.0*c/res.size(),s/c); }
}
