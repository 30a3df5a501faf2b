/*
 * TEST INFRASTRUCTURE — plain-C restatement of the MR_RL env hot path (batched, OpenMP).
 * Independent of the CUDA sources: written against the reference control flow
 *   MR_simulator.py:21-91 (Simulator), MR_env.py:70-201 (MR_Env) and scipy 1.18.1
 *   integrate/_ivp/rk.py (rk_step, RungeKutta._step_impl, RK45 tableau), common.py
 *   (norm, select_initial_step), base.py (OdeSolver.step).
 * Used (a) by tests for large-N parity where the Python oracle is too slow and (b) by bench.py as
 * the "C port" CPU baseline.  Pinned against the tests/golden npz files via tests/test_c_oracle.py.
 * Never linked into the product library.
 */
#include <math.h>
#include <pthread.h>
#include <unistd.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define N_STAGES 6
static const double C_B[N_STAGES] = {35.0 / 384, 0, 500.0 / 1113, 125.0 / 192, -2187.0 / 6784, 11.0 / 84};
static const double C_E[N_STAGES + 1] = {-71.0 / 57600, 0, 71.0 / 16695, -71.0 / 1920, 17253.0 / 339200, -22.0 / 525, 1.0 / 40};
#define SAFETY 0.9
#define MIN_FACTOR 0.2
#define MAX_FACTOR 10.0

typedef struct {
    double a0, sigma, dt, rtol, atol;
    int mism;
} sim_cfg;

typedef struct {           /* noise source: table stream or xorshift+Box-Muller */
    const double* z;
    int64_t cursor, len;
    uint64_t s0, s1;
    int use_rng, overflow;
    double spare;
    int has_spare;
} noise_src;

typedef struct {
    double y[2], f[2], h_abs, t, t_bound, sp[2];
    int counter, failed, attempts;
} env_t;

static uint64_t xs128p(noise_src* n) {
    uint64_t a = n->s0, b = n->s1;
    n->s0 = b;
    a ^= a << 23;
    n->s1 = a ^ b ^ (a >> 17) ^ (b >> 26);
    return n->s1 + b;
}

static double std_normal(noise_src* n) {
    if (!n->use_rng) {
        if (n->cursor >= n->len) { n->overflow = 1; n->cursor++; return 0.0; }
        return n->z[n->cursor++];
    }
    n->cursor++;
    if (n->has_spare) { n->has_spare = 0; return n->spare; }
    double u1 = ((xs128p(n) >> 11) + 1.0) * (1.0 / 9007199254740993.0);
    double u2 = (xs128p(n) >> 11) * (1.0 / 9007199254740992.0);
    double r = sqrt(-2.0 * log(u1)), a = 6.283185307179586 * u2;
    n->spare = r * sin(a);
    n->has_spare = 1;
    return r * cos(a);
}

static double gauss(noise_src* n, double mu, double sigma) { return mu + sigma * std_normal(n); }

/* Simulator.simulate (MR_simulator.py:58-88) */
static void simulate(const sim_cfg* c, const double act[2], noise_src* nz, double out[2], env_t* e) {
    double f = act[0], al = act[1];
    if (c->mism) {
        double a0 = c->a0 + (f / 4) * 0.8 + gauss(nz, 0, c->sigma / 4);
        out[0] = a0 * f * cos(al + 0.1) + gauss(nz, 0, c->sigma) + 0.2;
        out[1] = a0 * f * sin(al - 0.15) + gauss(nz, 0, c->sigma) - 0.1;
    } else {
        out[0] = c->a0 * f * cos(al) + gauss(nz, 0, c->sigma);
        out[1] = c->a0 * f * sin(al) + gauss(nz, 0, c->sigma);
    }
    e->sp[0] = out[0];
    e->sp[1] = out[1];
}

static double rms_norm2(double u, double v) { return sqrt(u * u + v * v) / sqrt(2.0); }

/* RK45.__init__ -> select_initial_step (common.py) */
static void make_integrator(env_t* e, const sim_cfg* c, const double act[2], noise_src* nz, double t0) {
    double f0[2], f1[2], scale[2];
    e->t = t0;
    e->t_bound = t0 + c->dt;
    simulate(c, act, nz, f0, e);
    e->f[0] = f0[0]; e->f[1] = f0[1];
    double interval = fabs(e->t_bound - t0);
    if (interval == 0.0) { e->h_abs = 0.0; return; }
    for (int i = 0; i < 2; i++) scale[i] = c->atol + fabs(e->y[i]) * c->rtol;
    double d0 = rms_norm2(e->y[0] / scale[0], e->y[1] / scale[1]);
    double d1 = rms_norm2(f0[0] / scale[0], f0[1] / scale[1]);
    double h0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : 0.01 * d0 / d1;
    if (interval < h0) h0 = interval;
    simulate(c, act, nz, f1, e);
    double d2 = rms_norm2((f1[0] - f0[0]) / scale[0], (f1[1] - f0[1]) / scale[1]) / h0;
    double h1;
    if (d1 <= 1e-15 && d2 <= 1e-15) h1 = (h0 * 1e-3 > 1e-6) ? h0 * 1e-3 : 1e-6;
    else h1 = pow(0.01 / (d2 > d1 ? d2 : d1), 1.0 / 5.0);
    double h = 100 * h0;
    if (h1 < h) h = h1;
    if (interval < h) h = interval;
    e->h_abs = h;
}

/* OdeSolver.step loop of Simulator.step (MR_simulator.py:42-43) + RungeKutta._step_impl */
static void integrate_to_bound(env_t* e, const sim_cfg* c, const double act[2], noise_src* nz) {
    double K[N_STAGES + 1][2];
    e->attempts = 0;
    while (!(e->t - e->t_bound >= 0)) {
        double t = e->t;
        double min_step = 10 * fabs(nextafter(t, INFINITY) - t);
        double h_abs = e->h_abs < min_step ? min_step : e->h_abs;
        int rejected = 0, accepted = 0;
        double t_new = t, y_new[2] = {0, 0};
        while (!accepted) {
            if (h_abs < min_step) { e->failed = 1; return; }
            t_new = t + h_abs;
            if (t_new - e->t_bound > 0) t_new = e->t_bound;
            double h = t_new - t;
            h_abs = fabs(h);
            K[0][0] = e->f[0]; K[0][1] = e->f[1];
            for (int s = 1; s < N_STAGES; s++) simulate(c, act, nz, K[s], e);   /* stage positions unused: RHS ignores y */
            for (int i = 0; i < 2; i++) {
                double acc = 0;
                for (int s = 0; s < N_STAGES; s++) acc += K[s][i] * C_B[s];
                y_new[i] = e->y[i] + h * acc;
            }
            simulate(c, act, nz, K[N_STAGES], e);
            e->attempts++;
            double errn[2];
            for (int i = 0; i < 2; i++) {
                double ya = fabs(e->y[i]), yb = fabs(y_new[i]);
                double scale = c->atol + (ya > yb ? ya : yb) * c->rtol;
                double acc = 0;
                for (int s = 0; s <= N_STAGES; s++) acc += K[s][i] * C_E[s];
                errn[i] = acc * h / scale;
            }
            double en = rms_norm2(errn[0], errn[1]);
            if (en < 1) {
                double factor;
                if (en == 0) factor = MAX_FACTOR;
                else { double v = SAFETY * pow(en, -0.2); factor = v < MAX_FACTOR ? v : MAX_FACTOR; }
                if (rejected && !(factor < 1)) factor = 1;
                h_abs *= factor;
                accepted = 1;
            } else {
                double v = SAFETY * pow(en, -0.2);
                h_abs *= (v > MIN_FACTOR ? v : MIN_FACTOR);
                rejected = 1;
            }
        }
        e->t = t_new;
        e->y[0] = y_new[0]; e->y[1] = y_new[1];
        e->f[0] = K[N_STAGES][0]; e->f[1] = K[N_STAGES][1];
        e->h_abs = h_abs;
    }
}

static int episode_over(const env_t* e, double* d_out) {
    double d = sqrt(e->y[0] * e->y[0] + e->y[1] * e->y[1]);     /* np.linalg.norm(goal - cur), goal = 0 */
    *d_out = d;
    int inside = e->y[0] >= -5000 && e->y[0] <= 5000 && e->y[1] >= -5000 && e->y[1] <= 5000 && d >= 0 && d <= 80000;
    return (!inside || e->counter > 50) || d < 30;             /* MR_env.py:136-152 */
}

/*
 * reset + T steps for n envs.  init [n][2]; actions [T][n][2]; z [n][zlen] (row per env) or NULL
 * to use the internal generator (seed).  Outputs may be NULL.  Returns the number of envs whose
 * solver failed or whose noise stream overflowed.  Envs are split over host threads (pthreads).
 */
typedef struct {
    int n, T, mism, mism_at_reset, auto_reset;
    const double *init, *actions, *z;
    double sigma, a0;
    int64_t zlen;
    uint64_t seed;
    double* pos_out; uint8_t* done_out; int64_t* cursor_out; int32_t* attempts_out; double* final_out;
    int i0, i1, bad;
} job_t;

static void* run_range(void* arg) {
    job_t* j = (job_t*)arg;
    const int n = j->n, T = j->T;
    int bad = 0;
    for (int i = j->i0; i < j->i1; i++) {
        sim_cfg c = {j->a0, j->sigma, 0.030, 0.030 / 100, 1e-4, j->mism_at_reset};
        noise_src nz;
        memset(&nz, 0, sizeof(nz));
        if (j->z) { nz.z = j->z + (int64_t)i * j->zlen; nz.len = j->zlen; }
        else { nz.use_rng = 1; nz.s0 = j->seed * 0x9E3779B97F4A7C15ull + (uint64_t)i + 1; nz.s1 = (uint64_t)i * 0xBF58476D1CE4E5B9ull + 0x94D049BB133111EBull; }
        env_t e;
        memset(&e, 0, sizeof(e));
        const double zero_act[2] = {0, 0};
        e.y[0] = j->init[2 * i]; e.y[1] = j->init[2 * i + 1];
        make_integrator(&e, &c, zero_act, &nz, 0.0);            /* MR_env.py:181, old mismatch flag */
        c.mism = j->mism;                                       /* MR_env.py:183 */
        for (int k = 0; k < T; k++) {
            const double* act = j->actions + ((int64_t)k * n + i) * 2;
            e.counter++;
            integrate_to_bound(&e, &c, act, &nz);
            if (e.failed) break;
            make_integrator(&e, &c, act, &nz, e.t);             /* MR_simulator.py:46-50 */
            double d;
            int done = episode_over(&e, &d);
            if (j->pos_out) { j->pos_out[((int64_t)k * n + i) * 2] = e.y[0]; j->pos_out[((int64_t)k * n + i) * 2 + 1] = e.y[1]; }
            if (j->done_out) j->done_out[(int64_t)k * n + i] = (uint8_t)done;
            if (j->attempts_out) j->attempts_out[(int64_t)k * n + i] = e.attempts;
            if (done && j->auto_reset) {
                e.y[0] = j->init[2 * i]; e.y[1] = j->init[2 * i + 1];
                e.counter = 0;
                make_integrator(&e, &c, zero_act, &nz, 0.0);
            }
        }
        if (j->cursor_out) j->cursor_out[i] = nz.cursor;
        if (j->final_out) { double* o = j->final_out + 5 * (int64_t)i; o[0] = e.y[0]; o[1] = e.y[1]; o[2] = e.f[0]; o[3] = e.f[1]; o[4] = e.h_abs; }
        bad += e.failed || nz.overflow;
    }
    j->bad = bad;
    return NULL;
}

int mr_oracle_max_threads(void) {
    const char* env = getenv("MR_ORACLE_THREADS");
    long n = env ? atol(env) : sysconf(_SC_NPROCESSORS_ONLN);
    if (n < 1) n = 1;
    if (n > 256) n = 256;
    return (int)n;
}

int mr_oracle_rollout(int n, int T, const double* init, const double* actions, double sigma, double a0, int mism,
                      int mism_at_reset, const double* z, int64_t zlen, uint64_t seed, int auto_reset,
                      double* pos_out /*[T][n][2]*/, uint8_t* done_out /*[T][n]*/, int64_t* cursor_out /*[n]*/,
                      int32_t* attempts_out /*[T][n]*/, double* final_out /*[n][5]: x y fx fy h*/) {
    int nt = mr_oracle_max_threads();
    if (nt > (n + 63) / 64) nt = (n + 63) / 64;
    if (nt < 1) nt = 1;
    job_t jobs[256];
    pthread_t th[256];
    for (int t = 0; t < nt; t++) {
        job_t j = {n, T, mism, mism_at_reset, auto_reset, init, actions, z, sigma, a0, zlen, seed,
                   pos_out, done_out, cursor_out, attempts_out, final_out,
                   (int)((int64_t)n * t / nt), (int)((int64_t)n * (t + 1) / nt), 0};
        jobs[t] = j;
    }
    for (int t = 1; t < nt; t++) pthread_create(&th[t], NULL, run_range, &jobs[t]);
    run_range(&jobs[0]);
    int bad = jobs[0].bad;
    for (int t = 1; t < nt; t++) { pthread_join(th[t], NULL); bad += jobs[t].bad; }
    return bad;
}
