/* A plain-C client of the drop-in boundary (include/wae.h): what a binding in any language does, without Python or torch.
 *
 *   gcc -std=c99 -Iinclude examples/c_client.c -Lweb-audio-api-rs_b200 -lwae_b200 -Wl,-rpath,$PWD/web-audio-api-rs_b200 -lm -o c_client
 *
 * Builds N OfflineAudioContexts (Oscillator -> BiquadFilter(lowpass) -> Gain -> destination, the graph of tests/offline.rs with a
 * gain) WITHOUT a device (graph construction is host work), asks the library what it would lower them to (wae_batch_plan, host only),
 * then creates the engine, renders them with ONE wae_render_batch call into host memory and prints the RMS of every render.
 * Exit code: 0 = rendered, 2 = no usable CUDA device (the library has no CPU fallback and says so), 1 = any other error. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "wae.h"

#define CHECK(call)                                                                   \
    do {                                                                              \
        wae_status st_ = (call);                                                      \
        if (st_ != WAE_OK) {                                                          \
            fprintf(stderr, "%s -> %d: %s\n", #call, (int)st_, wae_last_error());     \
            return st_ == WAE_NO_DEVICE ? 2 : 1;                                      \
        }                                                                             \
    } while (0)

int main(int argc, char** argv) {
    const uint32_t n_graphs = argc > 1 ? (uint32_t)atoi(argv[1]) : 4;
    const uint64_t length = 48000; /* 1 s at 48 kHz */
    wae_graph** graphs = (wae_graph**)calloc(n_graphs, sizeof(wae_graph*));
    for (uint32_t g = 0; g < n_graphs; g++) {
        CHECK(wae_graph_create(NULL, 2, length, 48000.f, &graphs[g]));
        wae_oscillator_options osc;
        memset(&osc, 0, sizeof osc);
        osc.type = WAE_OSC_SAWTOOTH;
        osc.frequency = 110.f * (float)(g + 1);
        wae_biquad_options bq;
        memset(&bq, 0, sizeof bq); /* channel_config.count == 0: the node's default config */
        bq.type = WAE_BIQUAD_LOWPASS;
        bq.frequency = 900.f;
        bq.q = 2.f;
        wae_gain_options gain;
        memset(&gain, 0, sizeof gain);
        gain.gain = 0.5f;
        wae_node_id osc_id, bq_id, gain_id;
        CHECK(wae_create_oscillator(graphs[g], &osc, &osc_id));
        CHECK(wae_create_biquad_filter(graphs[g], &bq, &bq_id));
        CHECK(wae_create_gain(graphs[g], &gain, &gain_id));
        CHECK(wae_connect(graphs[g], osc_id, 0, bq_id, 0));
        CHECK(wae_connect(graphs[g], bq_id, 0, gain_id, 0));
        CHECK(wae_connect(graphs[g], gain_id, 0, /* destination */ 0, 0));
        /* an envelope on the gain: AudioParam::linear_ramp_to_value_at_time(0, 1.0) */
        wae_param_event ev;
        memset(&ev, 0, sizeof ev);
        ev.type = WAE_EVENT_LINEAR_RAMP_TO_VALUE_AT_TIME;
        ev.value = 0.f;
        ev.time = 1.0;
        CHECK(wae_param_event_push(graphs[g], gain_id, 0, &ev));
        CHECK(wae_source_start(graphs[g], osc_id, 0.0, 0.0, 0.0));
    }

    wae_plan_info plan;
    CHECK(wae_batch_plan(graphs, n_graphs, &plan));
    printf("plan: %u stage(s) per chunk [%s], %llu chunk(s) of %llu frames\n", plan.stages, plan.stage_kinds, (unsigned long long)plan.chunks,
           (unsigned long long)plan.chunk_frames);

    wae_engine* engine = NULL;
    CHECK(wae_engine_create(0, &engine));
    float* pcm = (float*)malloc((size_t)n_graphs * 2 * length * sizeof(float)); /* [graph][channel][frame] */
    CHECK(wae_render_batch(engine, graphs, n_graphs, pcm, WAE_RENDER_OUT_HOST));
    for (uint32_t g = 0; g < n_graphs; g++) {
        double acc = 0.;
        const float* left = pcm + (size_t)g * 2 * length;
        for (uint64_t i = 0; i < length; i++) acc += (double)left[i] * left[i];
        printf("graph %u: rms %.6f\n", g, sqrt(acc / (double)length));
    }
    for (uint32_t g = 0; g < n_graphs; g++) CHECK(wae_graph_destroy(graphs[g]));
    CHECK(wae_engine_destroy(engine));
    free(pcm);
    free(graphs);
    return 0;
}
