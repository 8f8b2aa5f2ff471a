// PCA scatter of the MLP2 outputs (reference components/PCAPlot.vue): one point per (layer, token); points of a token are
// joined into its trajectory through the layers; colour = layer, the selected token is emphasised.
import {Component, h, panel} from "./base.js";
import {state, set, flat, shapeOf} from "../store.js";
import {pca2} from "../draw.js";
export class PCAPlot extends Component {
  constructor() { super(); this.cv = h("canvas", {width:520, height:340}); this.info = h("small", {});
    this.el = panel("MLP2 PCA trajectories", this.cv, this.info); this.watch("data", "tokens"); this.pts = []; this.cv.onclick = ev => this.pick(ev); }
  collect() {
    const pts = [];
    for (const k of Object.keys(state.updates)) { const [flag, layer] = k.split(":").map(Number); if (flag !== 7) continue;
      let tok = 0; state.updates[k].forEach(e => { const shp = shapeOf(e), d = shp[shp.length - 1] || 1, f = flat(e.result);
        for (let i = 0; i + d <= f.length; i += d) pts.push({layer, tok: tok++, v: f.slice(i, i + d)}); }); }
    return pts;
  }
  render() {
    const ctx = this.cv.getContext("2d"), W = this.cv.width, H = this.cv.height; ctx.clearRect(0, 0, W, H);
    const pts = this.collect().filter(p => p.v.length > 1); this.pts = [];
    if (pts.length < 3) { this.info.textContent = "enable 'MLP2 PCA trajectory' to collect points"; return; }
    const d = Math.min(...pts.map(p => p.v.length)), {points:P, explained} = pca2(pts.map(p => p.v.slice(0, d)));
    const xs = P.map(p => p[0]), ys = P.map(p => p[1]), x0 = Math.min(...xs), x1 = Math.max(...xs), y0 = Math.min(...ys), y1 = Math.max(...ys);
    const X = v => 24 + (v - x0) / (x1 - x0 + 1e-9) * (W - 48), Y = v => H - 24 - (v - y0) / (y1 - y0 + 1e-9) * (H - 48);
    const byTok = new Map(); pts.forEach((p, i) => { (byTok.get(p.tok) || byTok.set(p.tok, []).get(p.tok)).push([p.layer, i]); });
    byTok.forEach((arr, tok) => { arr.sort((a, b) => a[0] - b[0]); ctx.strokeStyle = tok === state.selToken ? "#76b900" : "#3a4658"; ctx.lineWidth = tok === state.selToken ? 2 : 1; ctx.beginPath();
      arr.forEach(([, i], j) => j ? ctx.lineTo(X(P[i][0]), Y(P[i][1])) : ctx.moveTo(X(P[i][0]), Y(P[i][1]))); ctx.stroke(); });
    pts.forEach((p, i) => { const x = X(P[i][0]), y = Y(P[i][1]); this.pts.push([x, y, p.tok]); ctx.fillStyle = `hsl(${(p.layer * 47) % 360} 80% 60%)`; ctx.beginPath();
      ctx.arc(x, y, p.tok === state.selToken ? 5 : 3, 0, 6.3); ctx.fill(); });
    this.info.textContent = `${pts.length} points, explained variance ${(100 * explained[0]).toFixed(1)} % / ${(100 * explained[1]).toFixed(1)} %  (colour = layer, click a point to select its token)`;
  }
  pick(ev) { const b = this.cv.getBoundingClientRect(), x = (ev.clientX - b.left) * this.cv.width / b.width, y = (ev.clientY - b.top) * this.cv.height / b.height;
    let best = -1, bd = 144; this.pts.forEach(([px, py, tok]) => { const d = (px - x) ** 2 + (py - y) ** 2; if (d < bd) { bd = d; best = tok; } }); if (best >= 0) set({selToken:best}, "tokens"); }
}
