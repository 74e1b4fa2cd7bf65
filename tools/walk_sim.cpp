// walk_sim.cpp -- CPU simulation of the warp-level structure of the Barnes-Hut walk (design tool, not on the product path).
//
// Builds the reference's one-particle-per-leaf octree for a uniform or Gaussian sphere, groups the key-sorted particles
// in warps of 32 and replays the reference's acceptance rule (F/gravOctree_Single.jl:265) for
//   * the shared masked depth-first walk (one cell per warp iteration, lane masks),
//   * a hybrid: cells that <= T lanes must open are deferred to per-lane private walks (drained when >= DHI lanes have
//     work, until < DLO), and T = 32, DHI = 1: fully independent per-lane walks.
// Prints warp iterations per group, lane utilisation and the number of distinct 128-byte lines (two 64-byte node records
// per line, BFS order) touched per private iteration.  These numbers motivated the pair queue of walk_pairs_kernel
// (DESIGN.md section 4): 1 843 shared iterations serve 901 visits per lane (packing 0.49); independent walks need 975
// iterations but touch ~20 lines per load instruction; sparse cells are half of the iterations and a seventh of the visits.
//
// build / run:  g++ -O2 -o /tmp/walk_sim tools/walk_sim.cpp && /tmp/walk_sim 1000000 12 24 12 [dist: 0 uniform, 1 gaussian] [hilbert: 0/1]
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <algorithm>
#include <random>
#include <cstdint>
#include <set>
#include <deque>
using namespace std;
struct Node { double c[3], L; double com[3], M; int child[8]; int nch; int part; double lo[3], hi[3]; int bfs; };
vector<Node> nodes; vector<double> X, Y, Z, H;
int build(vector<int>& idx, double cx, double cy, double cz, double L) {
    int id = nodes.size(); nodes.push_back(Node());
    { Node& n = nodes[id]; n.c[0]=cx; n.c[1]=cy; n.c[2]=cz; n.L=L; n.nch=0; n.part=-1;
      n.lo[0]=cx-L; n.lo[1]=cy-L; n.lo[2]=cz-L; n.hi[0]=cx+L; n.hi[1]=cy+L; n.hi[2]=cz+L; }
    if (idx.size()==1) { Node& n=nodes[id]; n.part=idx[0]; n.com[0]=X[idx[0]]; n.com[1]=Y[idx[0]]; n.com[2]=Z[idx[0]]; n.M=1; return id; }
    vector<int> sub[8];
    for (int i: idx) { int o=(X[i]-cx>0)+2*(Y[i]-cy>0)+4*(Z[i]-cz>0); sub[o].push_back(i); }
    double m=0, s[3]={0,0,0};
    for (int o=0;o<8;++o) if(!sub[o].empty()) {
        double h=L/2; int ch=build(sub[o], cx+((o&1)?h:-h), cy+((o&2)?h:-h), cz+((o&4)?h:-h), h);
        Node& n=nodes[id]; n.child[n.nch++]=ch; m+=nodes[ch].M; for(int k=0;k<3;++k) s[k]+=nodes[ch].M*nodes[ch].com[k];
    }
    Node& n=nodes[id]; n.M=m; for(int k=0;k<3;++k) n.com[k]=s[k]/m; return id;
}
const double theta=0.576;
inline bool accept(int i, const Node& n){ double dx=X[i]-n.com[0],dy=Y[i]-n.com[1],dz=Z[i]-n.com[2]; double d2=dx*dx+dy*dy+dz*dz; double s=2*n.L;
    bool acc = s*s/d2<theta*theta; if(acc){ double e2=0; double p[3]={X[i],Y[i],Z[i]}; for(int k=0;k<3;++k){double a=max(max(n.lo[k]-p[k],0.0),p[k]-n.hi[k]); e2+=a*a;} acc = H[i]*H[i]/e2<0.25; } return acc; }
int main(int argc,char**argv){
    int N=argc>1?atoi(argv[1]):100000; int T=argc>2?atoi(argv[2]):8; int DHI=argc>3?atoi(argv[3]):24; int DLO=argc>4?atoi(argv[4]):12; int dist=argc>5?atoi(argv[5]):0;
    const int G=32;
    mt19937_64 rng(1); uniform_real_distribution<double> U(-1,1); normal_distribution<double> Nn(0,0.3);
    if(dist==0) while((int)X.size()<N){double x=U(rng),y=U(rng),z=U(rng); if(x*x+y*y+z*z<=1){X.push_back(x);Y.push_back(y);Z.push_back(z);}}
    else for(int i=0;i<N;++i){X.push_back(Nn(rng));Y.push_back(Nn(rng));Z.push_back(Nn(rng));}
    double l=0; for(int i=0;i<N;++i) l=max(l,max(fabs(X[i]),max(fabs(Y[i]),fabs(Z[i]))));
    H.assign(N, 0.5*cbrt(50.0*3/(4*M_PI)/ (N/(4*M_PI/3))));
    vector<int> all(N); for(int i=0;i<N;++i) all[i]=i;
    nodes.reserve(2*N); int root=build(all,0,0,0,l);
    { deque<int> q{root}; int k=0; while(!q.empty()){int n=q.front(); q.pop_front(); nodes[n].bfs=k++; for(int c=0;c<nodes[n].nch;++c) q.push_back(nodes[n].child[c]);} }
    vector<int> order; { vector<int> st{root}; while(!st.empty()){int n=st.back(); st.pop_back(); if(nodes[n].part>=0) order.push_back(nodes[n].part); else for(int c=nodes[n].nch-1;c>=0;--c) st.push_back(nodes[n].child[c]);} }
    if (argc > 6 && atoi(argv[6]) == 1) {
        // regroup the targets along a 3-D Hilbert curve (10 bits per axis) instead of the tree's Z-order: how much of the
        // packing loss is due to the jumps of the Z-order?
        auto hilbert = [&](int i) {
            unsigned X3[3] = {(unsigned)((X[i] + l) / (2 * l) * 1023.999), (unsigned)((Y[i] + l) / (2 * l) * 1023.999), (unsigned)((Z[i] + l) / (2 * l) * 1023.999)};
            const int b = 10; unsigned M = 1u << (b - 1), P, Q, t;
            for (Q = M; Q > 1; Q >>= 1) { P = Q - 1; for (int k = 0; k < 3; ++k) { if (X3[k] & Q) X3[0] ^= P; else { t = (X3[0] ^ X3[k]) & P; X3[0] ^= t; X3[k] ^= t; } } }
            for (int k = 1; k < 3; ++k) X3[k] ^= X3[k - 1];
            t = 0; for (Q = M; Q > 1; Q >>= 1) if (X3[2] & Q) t ^= Q - 1;
            for (int k = 0; k < 3; ++k) X3[k] ^= t;
            unsigned long long h = 0;
            for (int bit = b - 1; bit >= 0; --bit) for (int k = 0; k < 3; ++k) h = (h << 1) | ((X3[k] >> bit) & 1u);
            return h;
        };
        vector<pair<unsigned long long,int>> hk(N);
        for (int s = 0; s < N; ++s) hk[s] = {hilbert(order[s]), order[s]};
        sort(hk.begin(), hk.end());
        for (int s = 0; s < N; ++s) order[s] = hk[s].second;
        printf("targets regrouped in Hilbert order\n");
    }
    if(dist==1){ // h from local density estimate: crude, use distance to 50th in key order window
        for(int s=0;s<N;++s){int i=order[s]; int a=max(0,s-25), b=min(N-1,s+25); double m=0; for(int t=a;t<=b;++t){int j=order[t]; double d=hypot(hypot(X[i]-X[j],Y[i]-Y[j]),Z[i]-Z[j]); m=max(m,d);} H[i]=0.5*m*0.6;} }
    long dense_visits=0, priv_iters=0, priv_lanevisits=0, lanevisits_dense=0, lines=0, drains=0, maxq=0;
    int ngroups=0;
    for (int g0=0; g0+G<=N; g0+= G*37) {
        ++ngroups;
        struct E{int n; unsigned m;}; vector<E> st; st.push_back({root,0xffffffffu});
        vector<vector<int>> q(G);           // per-lane deferred items: node to open
        vector<vector<int>> cur(G);         // per-lane DFS stack of the item in progress (sim of the stackless cursor)
        auto nonempty=[&](){int c=0; for(int l=0;l<G;++l) if(!q[l].empty()||!cur[l].empty()) ++c; return c;};
        auto drain=[&](int lo){ ++drains;
            while(true){ int act=nonempty(); if(act==0|| act<lo) break;
                ++priv_iters; set<int> ln;
                for(int l=0;l<G;++l){ if(cur[l].empty()){ if(q[l].empty()) continue; int P=q[l].back(); q[l].pop_back(); for(int c=nodes[P].nch-1;c>=0;--c) cur[l].push_back(nodes[P].child[c]); }
                    int n=cur[l].back(); cur[l].pop_back(); ++priv_lanevisits; ln.insert(nodes[n].bfs/2);
                    if(nodes[n].part<0 && !accept(order[g0+l],nodes[n])) for(int c=nodes[n].nch-1;c>=0;--c) cur[l].push_back(nodes[n].child[c]); }
                lines+=ln.size(); } };
        while(!st.empty()){ E e=st.back(); st.pop_back(); Node& P=nodes[e.n];
            int pc=__builtin_popcount(e.m);
            if(pc<=T){ for(int l=0;l<G;++l) if((e.m>>l)&1){ q[l].push_back(e.n); maxq=max<long>(maxq,q[l].size()); }
                if(nonempty()>=DHI) drain(DLO); continue; }
            for(int c=0;c<P.nch;++c){ Node& n=nodes[P.child[c]]; ++dense_visits; lanevisits_dense+=pc;
                if(n.part>=0) continue; unsigned om=0;
                for(int l=0;l<G;++l) if((e.m>>l)&1) if(!accept(order[g0+l],n)) om|=1u<<l;
                if(om) st.push_back({P.child[c],om}); } }
        drain(0);
    }
    printf("N=%d T=%d DHI=%d DLO=%d: per group: dense warp-visits %.0f (lane-visits/target %.0f)  private iterations %.0f (lane-visits/target %.0f, util %.2f, lines/iter %.1f)  drains %.1f maxq %ld  => warp-iterations %.0f (T = 0 gives the pure shared walk)\n",
      N,T,DHI,DLO,(double)dense_visits/ngroups,(double)lanevisits_dense/ngroups/G,(double)priv_iters/ngroups,(double)priv_lanevisits/ngroups/G,(double)priv_lanevisits/(priv_iters*32.0),(double)lines/priv_iters,(double)drains/ngroups,maxq,(double)(dense_visits+priv_iters)/ngroups);
}
