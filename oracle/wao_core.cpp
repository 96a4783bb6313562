// ORACLE — TEST INFRASTRUCTURE ONLY (see wao_core.h).
// Restates src/render/quantum.rs:274-586 (up/down-mix + add) and src/render/graph.rs:233-591
// (edges, topological order with cycle handling, per-quantum render loop).
#include "wao_core.h"

namespace wao {

// src/render/quantum.rs:285-505
void Quantum::mix_inner(int computed, int interp) {
    Channel silence = ch[0].silence();
    int cur = number_of_channels();
    auto fill_or_truncate = [&]() {
        for (int i = number_of_channels(); i < computed; i++) ch.push_back(silence);
        ch.resize(computed, silence);
    };
    if (interp == DISCRETE || cur > 6 || computed > 6) {
        fill_or_truncate();
        return;
    }
    auto key = [](int a, int b) { return a * 16 + b; };
    switch (key(cur, computed)) {
        // ---- up-mix (quantum.rs:301-377)
        case 1 * 16 + 2:
            ch.push_back(ch[0]);
            break;
        case 1 * 16 + 4:
            ch.push_back(ch[0]);
            ch.push_back(silence);
            ch.push_back(silence);
            break;
        case 1 * 16 + 6: {
            Channel main = ch[0];
            ch[0] = silence;
            ch.push_back(silence);
            ch.push_back(main);
            ch.push_back(silence);
            ch.push_back(silence);
            ch.push_back(silence);
            break;
        }
        case 2 * 16 + 4:
            ch.push_back(silence);
            ch.push_back(silence);
            break;
        case 2 * 16 + 6:
            ch.push_back(silence);
            ch.push_back(silence);
            ch.push_back(silence);
            ch.push_back(silence);
            break;
        case 4 * 16 + 5: {
            // quantum.rs:358-366: sl = replace(ch[2], silence); sr = replace(ch[3], sl); push(sr)
            Channel sl = ch[2];
            ch[2] = silence;
            Channel sr = ch[3];
            ch[3] = sl;
            ch.push_back(sr);
            break;
        }
        case 4 * 16 + 6: {
            Channel sl = ch[2];
            ch[2] = silence;
            Channel sr = ch[3];
            ch[3] = silence;
            ch.push_back(sl);
            ch.push_back(sr);
            break;
        }
        // ---- down-mix (quantum.rs:382-490)
        case 2 * 16 + 1: {
            Channel right = ch[1];
            float* l = ch[0].make_mut();
            const float* r = right.data();
            for (int i = 0; i < RQ; i++) l[i] = 0.5f * (l[i] + r[i]);
            ch.resize(1, silence);
            break;
        }
        case 4 * 16 + 1: {
            Channel right = ch[1], sl = ch[2], sr = ch[3];
            float* l = ch[0].make_mut();
            for (int i = 0; i < RQ; i++) l[i] = 0.25f * (l[i] + right.data()[i] + sl.data()[i] + sr.data()[i]);
            ch.resize(1, silence);
            break;
        }
        case 6 * 16 + 1: {
            Channel right = ch[1], center = ch[2], sl = ch[4], sr = ch[5];
            float sqrt05 = std::sqrt(0.5f);
            float* l = ch[0].make_mut();
            for (int i = 0; i < RQ; i++)
                l[i] = std::fma(sqrt05, l[i] + right.data()[i], std::fma(0.5f, sl.data()[i] + sr.data()[i], center.data()[i]));
            ch.resize(1, silence);
            break;
        }
        case 4 * 16 + 2: {
            Channel sl = ch[2], sr = ch[3];
            float* l = ch[0].make_mut();
            for (int i = 0; i < RQ; i++) l[i] = 0.5f * (l[i] + sl.data()[i]);
            float* r = ch[1].make_mut();
            for (int i = 0; i < RQ; i++) r[i] = 0.5f * (r[i] + sr.data()[i]);
            ch.resize(2, silence);
            break;
        }
        case 6 * 16 + 2: {
            Channel center = ch[2], sl = ch[4], sr = ch[5];
            float sqrt05 = std::sqrt(0.5f);
            float* l = ch[0].make_mut();
            for (int i = 0; i < RQ; i++) l[i] += sqrt05 * (center.data()[i] + sl.data()[i]);
            float* r = ch[1].make_mut();
            for (int i = 0; i < RQ; i++) r[i] += sqrt05 * (center.data()[i] + sr.data()[i]);
            ch.resize(2, silence);
            break;
        }
        case 6 * 16 + 4: {
            // quantum.rs:474-489: swap_remove(3) then swap_remove(2) -> [L, R, SL, SR], centre kept aside
            Channel center = ch[2];
            Channel sl = ch[4], sr = ch[5];
            ch[2] = sl;
            ch[3] = sr;
            ch.resize(4, silence);
            float sqrt05 = std::sqrt(0.5f);
            float* l = ch[0].make_mut();
            for (int i = 0; i < RQ; i++) l[i] += sqrt05 * center.data()[i];
            float* r = ch[1].make_mut();
            for (int i = 0; i < RQ; i++) r[i] += sqrt05 * center.data()[i];
            break;
        }
        default:
            fill_or_truncate();
    }
}

// src/render/quantum.rs:532-569
void Quantum::add(const Quantum& other, const ChannelConfig& cfg) {
    int channels_self = number_of_channels();
    int channels_other = other.number_of_channels();
    int max_channels = std::max(channels_self, channels_other);
    int new_channels;
    switch (cfg.mode) {
        case MODE_MAX: new_channels = max_channels; break;
        case MODE_EXPLICIT: new_channels = cfg.count; break;
        default: new_channels = std::min(max_channels, cfg.count);
    }
    if (cfg.interp == SPEAKERS && all_channels_identical() && other.all_channels_identical()) {
        ch.resize(1, ch[0]);
        ch[0].add(other.ch[0]);
        mix(new_channels, cfg.interp);
        return;
    }
    mix(new_channels, cfg.interp);
    Quantum other_mixed = other;
    other_mixed.mix(new_channels, cfg.interp);
    for (int i = 0; i < new_channels; i++) ch[i].add(other_mixed.ch[i]);
}

// ---- Graph -----------------------------------------------------------------------------------------

// src/render/graph.rs:233-267
void Graph::add_node(uint32_t id, std::unique_ptr<Processor> p, int n_in, int n_out, ChannelConfig cfg) {
    auto n = std::make_unique<Node>();
    n->processor = std::move(p);
    Channel silence(&alloc.zeroes, &alloc);
    for (int i = 0; i < n_in; i++) n->inputs.emplace_back(silence);
    for (int i = 0; i < n_out; i++) n->outputs.emplace_back(silence);
    n->cfg = cfg;
    nodes[id] = std::move(n);
    ordered.push_back(id);
}

// src/render/graph.rs:269-280
void Graph::add_edge(uint32_t src, int out, uint32_t dst, int in) {
    nodes.at(src)->outgoing.push_back(Edge{out, dst, in});
    ordered.clear();
}

void Graph::remove_edges_from(uint32_t src) {
    // hidden param edges are never created from user nodes, so clearing everything mirrors
    // AudioNode::disconnect() (src/node/audio_node.rs) for non-param sources
    nodes.at(src)->outgoing.clear();
    ordered.clear();
}

static bool contains(const std::vector<uint32_t>& v, uint32_t x) { return std::find(v.begin(), v.end(), x) != v.end(); }

// src/render/graph.rs:331-403
bool Graph::visit(uint32_t node_id) {
    auto it = std::find(marked_temp.begin(), marked_temp.end(), node_id);
    if (it != marked_temp.end()) {
        // part of a cycle: look for a cycle breaker among the nodes of the cycle
        for (auto jt = it; jt != marked_temp.end(); ++jt) {
            if (nodes.at(*jt)->cycle_breaker) {
                cycle_breakers.push_back(*jt);
                return true;
            }
        }
        in_cycle.insert(in_cycle.end(), it, marked_temp.end());
        return false;
    }
    if (contains(marked, node_id)) return false;
    marked.push_back(node_id);
    marked_temp.push_back(node_id);
    // note: iterate over a copy of the ids, the recursion never mutates edges
    const auto& edges = nodes.at(node_id)->outgoing;
    for (size_t i = 0; i < edges.size(); i++) {
        if (nodes.find(edges[i].other_id) == nodes.end()) continue;
        if (visit(edges[i].other_id)) return true;
    }
    ordered.push_back(node_id);
    marked_temp.erase(std::remove(marked_temp.begin(), marked_temp.end(), node_id), marked_temp.end());
    return false;
}

// src/render/graph.rs:418-487
void Graph::order_nodes() {
    for (;;) {
        ordered.clear();
        marked.clear();
        marked_temp.clear();
        in_cycle.clear();
        cycle_breakers.clear();
        bool applied = false;
        for (auto& kv : nodes) {
            applied = visit(kv.first);
            if (applied) break;
        }
        if (applied) {
            for (uint32_t id : cycle_breakers) nodes.at(id)->outgoing.clear();
            continue;
        }
        break;
    }
    ordered.erase(std::remove_if(ordered.begin(), ordered.end(), [&](uint32_t o) { return contains(in_cycle, o); }),
                  ordered.end());
    std::reverse(ordered.begin(), ordered.end());
}

// src/render/graph.rs:490-591 (node lifecycle / can_free is not restated: in an offline render the
// control handles outlive the render, so control_handle_dropped stays false — graph.rs:87-93)
const Quantum& Graph::render(const Scope& scope) {
    if (ordered.empty()) order_nodes();
    ParamValues params{this};
    for (uint32_t id : ordered) {
        Node* node = nodes.at(id).get();
        node->processor->process(node->inputs, node->outputs, params, scope);
        for (const Edge& e : node->outgoing) {
            if (e.other_index < 0) continue;  // hidden param edges (graph.rs:526-527)
            Node* dst = nodes.at(e.other_id).get();
            dst->has_inputs_connected = true;
            dst->inputs[e.other_index].add(node->outputs[e.self_index], dst->cfg);
        }
        for (auto& in : node->inputs) in.make_silent();
        node->has_inputs_connected = false;
    }
    return nodes.at(0)->outputs[0];
}

// src/render/processor.rs:231-247
ParamSlice ParamValues::get(uint32_t param_id) const {
    const Quantum& q = g->get(param_id)->outputs[0];
    if (q.single_valued) return ParamSlice{q.channel(0).data(), 1};
    return ParamSlice{q.channel(0).data(), RQ};
}

}  // namespace wao
