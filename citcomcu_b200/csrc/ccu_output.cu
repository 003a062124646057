// ccu_output.cu -- output staging: the fields the reference's output_velo_related (Output.c:54-170) prints leave
// the device asynchronously into pinned host buffers, and the ASCII files (.velo, .temp, .th_t / .th_b: same formats, so that
// utils/citcomcu_write_vtk and friends keep working) are written by a background thread while the solver goes on.
//
//   ccu_output_stage(ctx)            enqueue the device->host copies of T, V (and C with markers) behind everything already queued on the
//                                    context's stream; returns at once
//   ccu_output_write(ctx, ...)       hand the staged buffers to a writer thread (it waits for the copies itself)
//   ccu_output_wait(ctx)             join the writer; reports its error, if any
//
// At 8.5 M nodes the reference's fprintf loops dominate the wall time of an output step; here they cost the solver nothing.
#include "ccu_ctx.cuh"
#include <cstdio>
#include <string>
#include <thread>

struct CcuOutput
{
    float *T = nullptr, *V = nullptr, *C = nullptr, *hf = nullptr;      // pinned host copies
    size_t nno = 0;
    bool have_C = false, have_hf = false, staged = false;
    cudaEvent_t ready = nullptr;
    std::thread writer;
    std::string error;
};

static int out_init(ccu_ctx *c)
{
    if(c->out) return 0;
    CcuOutput *o = new CcuOutput();
    o->nno = (size_t)c->L[c->cfg.levmax].g.nno;
    CK(cudaMallocHost(&o->T, sizeof(float) * o->nno));
    CK(cudaMallocHost(&o->V, sizeof(float) * 3 * o->nno));
    CK(cudaMallocHost(&o->C, sizeof(float) * o->nno));
    CK(cudaMallocHost(&o->hf, sizeof(float) * o->nno));
    CK(cudaEventCreateWithFlags(&o->ready, cudaEventDisableTiming));
    c->out = o;
    return 0;
}
int ccu_output_wait(ccu_ctx *c)
{
    if(!c) FAIL("null context");
    CcuOutput *o = c->out;
    if(!o) return 0;
    if(o->writer.joinable()) o->writer.join();
    if(!o->error.empty()) { const std::string e = o->error; o->error.clear(); FAIL("output writer: " + e); }
    return 0;
}
void ccu_output_destroy(ccu_ctx *c)
{
    CcuOutput *o = c->out;
    if(!o) return;
    if(o->writer.joinable()) o->writer.join();
    cudaFreeHost(o->T); cudaFreeHost(o->V); cudaFreeHost(o->C); cudaFreeHost(o->hf);
    if(o->ready) cudaEventDestroy(o->ready);
    delete o;
    c->out = nullptr;
}
int ccu_output_stage(ccu_ctx *c)
{
    if(!c) FAIL("null context");
    if(out_init(c) || ccu_output_wait(c)) return 1;          // a previous write still owns the buffers
    CcuOutput *o = c->out;
    if(!c->T || !c->en.V || !c->en.have_v) FAIL("output_stage: temperature / velocity not resident");
    CK(cudaMemcpyAsync(o->T, c->T, sizeof(float) * o->nno, cudaMemcpyDeviceToHost, c->st));
    CK(cudaMemcpyAsync(o->V, c->en.V, sizeof(float) * 3 * o->nno, cudaMemcpyDeviceToHost, c->st));
    o->have_C = c->mk.ready;
    if(o->have_C) CK(cudaMemcpyAsync(o->C, c->mk.C, sizeof(float) * o->nno, cudaMemcpyDeviceToHost, c->st));
    o->have_hf = c->en.hf != nullptr;                         // nodal heat flux of the last ccu_heat_flux call, if any
    if(o->have_hf) CK(cudaMemcpyAsync(o->hf, c->en.hf, sizeof(float) * o->nno, cudaMemcpyDeviceToHost, c->st));
    CK(cudaEventRecord(o->ready, c->st));
    o->staged = true;
    return 0;
}
// prefix = E->control.data_file2, me = E->parallel.me, file_number = the step; timesteps = E->advection.timesteps
int ccu_output_write(ccu_ctx *c, const char *prefix, int me, int file_number, int timesteps, double elapsed_time, int composition)
{
    if(!c || !prefix) FAIL("output_write: null argument");
    CcuOutput *o = c->out;
    if(!o || !o->staged) FAIL("output_write: ccu_output_stage first");
    if(ccu_output_wait(c)) return 1;
    o->staged = false;
    const std::string pre(prefix);
    const int device = c->cfg.device;
    o->writer = std::thread([o, pre, me, file_number, timesteps, elapsed_time, composition, device]()
    {
        cudaSetDevice(device);
        if(cudaEventSynchronize(o->ready) != cudaSuccess) { o->error = "device copies failed"; return; }
        const size_t nno = o->nno;
        std::string buf;
        buf.reserve(64 * nno);
        char line[160];
        auto flush = [&](const std::string &name) -> bool
        {
            FILE *f = fopen(name.c_str(), "w");
            if(!f) { o->error = "cannot open " + name; return false; }
            const bool ok = fwrite(buf.data(), 1, buf.size(), f) == buf.size();
            fclose(f);
            buf.clear();
            if(!ok) o->error = "short write to " + name;
            return ok;
        };
        const std::string tail = "." + std::to_string(me) + "." + std::to_string(file_number);
        // .velo (Output.c:118-123)
        buf.append(line, (size_t)snprintf(line, sizeof line, "%6d %6d %.5e\n", (int)nno, timesteps, elapsed_time));
        for(size_t i = 0; i < nno; i++)
            buf.append(line, (size_t)snprintf(line, sizeof line, "%.6e %.6e %.6e\n", o->V[i], o->V[nno + i], o->V[2 * nno + i]));
        if(!flush(pre + ".velo" + tail)) return;
        // .temp (Output.c:93-116): T, Vz, the nodal heat flux column (E->heatflux_adv in the reference; here the nodal flux of the last
        // heat_flux call, 0 if none), C with a compositional field
        buf.append(line, (size_t)snprintf(line, sizeof line, "%6d %6d %.5e\n", (int)nno, timesteps, elapsed_time));
        for(size_t i = 0; i < nno; i++)
        {
            const float hf = o->have_hf ? o->hf[i] : 0.0f;
            if(composition && o->have_C) buf.append(line, (size_t)snprintf(line, sizeof line, "%.5e %.4e %.4e %.4e\n", o->T[i], o->V[2 * nno + i], hf, o->C[i]));
            else buf.append(line, (size_t)snprintf(line, sizeof line, "%.5e %.4e %.4e\n", o->T[i], o->V[2 * nno + i], hf));
        }
        flush(pre + ".temp" + tail);
    });
    return 0;
}
