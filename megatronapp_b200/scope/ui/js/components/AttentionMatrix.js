// Attention probabilities of one (layer, head): query x key heat-map with the token strip as axis read-out
// (reference components/AttentionMatrix.vue); "all heads" tiles every head of the layer.
import {Component, h, panel} from "./base.js";
import {state, set, latest, flat, shapeOf} from "../store.js";
import {drawMatrix, legend} from "../draw.js";

export class AttentionMatrix extends Component {
  constructor() {
    super(); this.cv = h("canvas", {width:360, height:360}); this.leg = h("canvas", {}); this.info = h("small", {});
    this.all = h("input", {type:"checkbox"}); this.all.onchange = () => this.render();
    this.el = panel("Attention matrix", this.cv, h("div", {class:"legend"}, "0", this.leg, "max", this.info),
                    h("div", {class:"chk"}, this.all, h("span", {}, "tile all heads")));
    legend(this.leg, true); this.watch("data", "view", "tokens");
  }
  render() {
    const e = latest(2, state.layer); if (!e) { this.cv.getContext("2d").clearRect(0, 0, 360, 360); this.info.textContent = "  no data"; return; }
    const shp = shapeOf(e), f = flat(e.result), sk = shp[shp.length - 1], sq = shp[shp.length - 2] || 1;
    const nh = shp.length >= 3 ? shp[shp.length - 3] : 1, per = nh * sq * sk, b = Math.min(state.batch, Math.max(0, Math.floor(f.length / per) - 1));
    if (state.meta.num_heads !== nh) { state.meta.num_heads = nh; set({}, "meta"); }
    const toks = state.promptTokens.concat(state.genTokens), name = i => toks[i] !== undefined ? toks[i] : i;
    if (this.all.checked) {
      const g = Math.ceil(Math.sqrt(nh));
      drawMatrix(this.cv, g * sq, g * sk, (r, c) => { const hh = Math.floor(r / sq) * g + Math.floor(c / sk); return hh < nh ? f[b * per + (hh * sq + r % sq) * sk + c % sk] || 0 : 0; },
                 {seq:true, label: (r, c) => `head ${Math.floor(r / sq) * g + Math.floor(c / sk)} q ${r % sq} k ${c % sk}`, onPick: (r, c) => set({head: Math.min(nh - 1, Math.floor(r / sq) * g + Math.floor(c / sk))})});
      this.info.textContent = `  layer ${state.layer}, ${nh} heads`;
    } else {
      const hd = Math.min(state.head, nh - 1);
      drawMatrix(this.cv, sq, sk, (r, c) => f[b * per + (hd * sq + r) * sk + c] || 0, {seq:true, highlightRow: sq > 1 ? state.selToken : -1,
                 label: (r, c) => `q ${name(r)} -> k ${name(c)}`, onPick: r => set({selToken:r}, "tokens")});
      this.info.textContent = `  layer ${state.layer} head ${hd} (${sq} x ${sk})`;
    }
  }
}
